"""AdamW on our own kernels with torch.optim.Optimizer's interface.

Drop-in for the `torch.optim.AdamW(param_groups, betas, eps)` the reference builds in
app/vjepa/utils.py:173-194: same param_groups / state layout (`step`, `exp_avg`, `exp_avg_sq`), so the
LR / WD schedulers, `adamw_logger`, `state_dict()` and reference checkpoints all keep working.
`_step_supports_amp_scaling` makes torch's GradScaler hand over `found_inf` / `grad_scale` as device
tensors instead of syncing the host: the skip-on-overflow decision is taken inside the kernel.

Fast path: when all parameters of a backbone live in a FlatParamStore and their gradients are the slices
of ONE flat gradient buffer (what our backward produces), the whole backbone is updated by a single
`vj_adamw_flat` launch; a per-64-element group table carries each tensor's (lr, weight_decay) group.
Anything else falls back to one `vj_adamw_step` launch per tensor.
"""
import ctypes

import torch

from . import _lib
from . import kernels as K


class FlatAdamW(torch.optim.Optimizer):
    _step_supports_amp_scaling = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._flat = {}   # id(store) -> dict(store, gid, m, v, step, key)

    # ------------------------------------------------------------------------------------------ flat path
    def _flat_plan(self):
        """Group parameters by FlatParamStore; returns (plans, leftovers)."""
        by_store, leftovers = {}, []
        for gi, group in enumerate(self.param_groups):
            for p in group['params']:
                store = getattr(p, "_vj_store", None)
                if store is None or gi > 3 or not store.owns(p):
                    leftovers.append((gi, p))
                else:
                    by_store.setdefault(id(store), (store, []))[1].append((gi, p))
        return by_store, leftovers

    def _flat_state(self, store, members):
        key = (store.flat.data_ptr(), store.total, tuple((gi, id(p)) for gi, p in members))
        st = self._flat.get(id(store))
        if st is not None and st["key"] == key:
            _, p0 = members[0]
            off0 = store.offsets[p0._vj_name][0]
            ea = self.state.get(p0, {}).get("exp_avg")
            if ea is not None and ea.data_ptr() == st["m"].data_ptr() + 4 * off0:
                return st          # state still aliases the flat moment buffers (not replaced by load_state_dict)
        dev = store.flat.device
        gid = torch.full((store.total // 64,), 255, dtype=torch.uint8)
        for gi, p in members:
            off, n, _ = store.offsets[p._vj_name]
            gid[off // 64:(off + n + 63) // 64] = gi if p.requires_grad else 255   # frozen (pos_embed): untouched
        m = torch.zeros(store.total, dtype=torch.float32, device=dev)
        v = torch.zeros(store.total, dtype=torch.float32, device=dev)
        step = torch.zeros((), dtype=torch.float32)
        for _, p in members:       # resume: carry per-tensor state (e.g. from load_state_dict) into the flat buffers
            old = self.state.get(p, {})
            if "step" in old:
                step = torch.as_tensor(float(old["step"]), dtype=torch.float32)
                break
        for gi, p in members:
            off, n, shape = store.offsets[p._vj_name]
            old = self.state.get(p, {})
            if "exp_avg" in old:
                m[off:off + n].view(shape).copy_(old["exp_avg"])
                v[off:off + n].view(shape).copy_(old["exp_avg_sq"])
            self.state[p] = {"step": step, "exp_avg": m[off:off + n].view(shape), "exp_avg_sq": v[off:off + n].view(shape)}
        st = dict(store=store, gid=gid.to(dev), m=m, v=v, step=step, key=key)
        self._flat[id(store)] = st
        return st

    @staticmethod
    def _flat_grad_base(store, members):
        """Device pointer of the flat gradient buffer if every member's .grad is its slice of one buffer."""
        base = None
        for _, p in members:
            g = p.grad
            if g is None:
                if p.requires_grad:
                    return None
                continue
            if g.dtype != torch.float32 or not g.is_contiguous():
                return None
            off = store.offsets[p._vj_name][0]
            b = g.data_ptr() - 4 * off
            if base is None:
                base = b
            elif b != base:
                return None
        return base

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

        by_store, leftovers = self._flat_plan()
        for store, members in by_store.values():
            base = self._flat_grad_base(store, members)
            # parameters of the store that are NOT optimised here (frozen ones) stay untouched: group id 255
            if base is None:
                leftovers.extend(members)
                continue
            st = self._flat_state(store, members)
            st["step"] += 1
            lr4 = (ctypes.c_float * 4)(*[float(g['lr']) for g in self.param_groups[:4]] + [0.0] * (4 - min(4, len(self.param_groups))))
            wd4 = (ctypes.c_float * 4)(*[float(g['weight_decay']) for g in self.param_groups[:4]] + [0.0] * (4 - min(4, len(self.param_groups))))
            beta1, beta2 = self.param_groups[0]['betas']
            _lib.call("vj_adamw_flat", store.flat.data_ptr(), base, st["m"].data_ptr(), st["v"].data_ptr(),
                      st["gid"].data_ptr(), store.total, ctypes.cast(lr4, ctypes.c_void_p), ctypes.cast(wd4, ctypes.c_void_p),
                      float(beta1), float(beta2), float(self.param_groups[0]['eps']), int(st["step"]),
                      K._p(inv_scale), K._p(found_inf), K._s())

        for gi, p in leftovers:
            if p.grad is None:
                continue
            group = self.param_groups[gi]
            beta1, beta2 = group['betas']
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
