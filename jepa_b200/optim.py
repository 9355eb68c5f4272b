"""AdamW on our own kernel (vj_adamw_step) with torch.optim.Optimizer's interface.

Drop-in for the `torch.optim.AdamW(param_groups, betas, eps)` the reference builds in
app/vjepa/utils.py:173-194: same param_groups / state layout (`step`, `exp_avg`, `exp_avg_sq`), so the
LR / WD schedulers, `adamw_logger`, `state_dict()` and reference checkpoints all keep working.
`_step_supports_amp_scaling` makes torch's GradScaler hand over `found_inf` / `grad_scale` as device
tensors instead of syncing the host: the skip-on-overflow decision is taken inside the kernel.
"""
import torch

from . import kernels as K


class FlatAdamW(torch.optim.Optimizer):
    _step_supports_amp_scaling = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        found_inf = getattr(self, "found_inf", None)
        grad_scale = getattr(self, "grad_scale", None)
        inv_scale = None
        if grad_scale is not None:
            inv_scale = grad_scale.double().reciprocal().float().reshape(1).contiguous()
        if found_inf is not None:
            found_inf = found_inf.float().reshape(1).contiguous()
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("FlatAdamW: parameters must be CUDA tensors (no CPU fallback)")
                state = self.state[p]
                if len(state) == 0:
                    state['step'] = torch.zeros((), dtype=torch.float32)
                    state['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state['step'] += 1
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                n = p.numel()
                if n % 4 != 0 or p.data_ptr() % 16 or g.data_ptr() % 16 or not p.is_contiguous():
                    raise RuntimeError(f"FlatAdamW: parameter of {n} elements is not 16-byte vectorisable; "
                                       "adopt the module into a FlatParamStore first")
                K.adamw_step(p, g, state['exp_avg'], state['exp_avg_sq'], group['lr'], beta1, beta2, group['eps'],
                             group['weight_decay'], int(state['step']), inv_scale, found_inf)
        return loss
