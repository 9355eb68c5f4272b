"""Host-logic dry run: the whole train-step orchestration (engine schedules, autograd glue, flat parameter
store, optimizer, EMA) executed on CPU tensors with every kernel ENTRY POINT replaced by a no-op, so shape /
stride / bookkeeping bugs in the Python layer are caught without a GPU.  No numerics are checked here
(outputs are uninitialised memory); parity lives in the -m gpu tests."""
import copy

import pytest
import torch

from common import C1
from parity_util import c1_masks


@pytest.fixture()
def dry(monkeypatch):
    from jepa_b200 import _lib, kernels, params
    calls = []

    def fake_call(name, *args):
        calls.append(name)
        return 0

    class FakeLib:
        @staticmethod
        def vj_layernorm_bwd_workspace(T, D):
            return 4 * D * 4 * 2

    monkeypatch.setattr(_lib, "call", fake_call)
    monkeypatch.setattr(_lib, "load", lambda: FakeLib)
    monkeypatch.setattr(kernels, "_s", lambda: 0)

    def chk(t, dtype=None, name="tensor"):
        assert t.is_contiguous(), name
        if dtype is not None:
            assert t.dtype == dtype, (name, t.dtype, dtype)
        return t

    monkeypatch.setattr(kernels, "_chk", chk)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True), raising=False)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: type("S", (), {"cuda_stream": 0})())
    return calls


def _adopt_on_cpu(monkeypatch):
    from jepa_b200 import params

    orig = params.FlatParamStore.adopt

    def adopt(self, module):
        named = [(n, p) for n, p in module.named_parameters()]
        if self._aliases(named):
            return self
        off, offsets = 0, {}
        for n, p in named:
            offsets[n] = (off, p.numel(), tuple(p.shape))
            off += (p.numel() + params.ALIGN - 1) // params.ALIGN * params.ALIGN
        flat = torch.zeros(off)
        with torch.no_grad():
            for n, p in named:
                o, cnt, shape = offsets[n]
                v = flat[o:o + cnt].view(shape)
                v.copy_(p.data)
                p.data = v
        self.flat, self.offsets, self.total, self._params = flat, offsets, off, named
        self.shadow = torch.empty(off, dtype=torch.bfloat16)
        return self

    monkeypatch.setattr(params.FlatParamStore, "adopt", adopt)


def test_full_step_orchestration_dry_run(dry, monkeypatch):
    _adopt_on_cpu(monkeypatch)
    from app.vjepa.utils import init_opt, init_video_model
    from jepa_b200 import step as vj
    enc, pred = init_video_model(device=torch.device("cpu"), patch_size=16, num_frames=C1["num_frames"], tubelet_size=2,
                                 model_name="vit_tiny", crop_size=C1["crop_size"], pred_depth=2, pred_embed_dim=384,
                                 uniform_power=True, use_mask_tokens=True, num_mask_tokens=2, use_sdpa=True)
    tgt = copy.deepcopy(enc)
    for p in tgt.parameters():
        p.requires_grad = False
    opt, _, sch, wds = init_opt(enc, pred, iterations_per_epoch=4, start_lr=1e-4, ref_lr=1e-3, warmup=1, num_epochs=1,
                                wd=0.04, final_wd=0.4)
    B = 2
    clips = torch.zeros(B, 3, C1["num_frames"], C1["crop_size"], C1["crop_size"])
    me, mp = c1_masks(B)
    for it in range(2):
        sch.step(); wds.step()
        h = vj.forward_target(tgt, clips, mp)
        assert [tuple(t.shape) for t in h] == [(B, m.shape[1], 192) for m in mp] and h[0].dtype == torch.float32
        z_enc = enc(clips, me)
        assert [tuple(t.shape) for t in z_enc] == [(B, m.shape[1], 192) for m in me]
        z = pred(z_enc, h, me, mp)
        assert [tuple(t.shape) for t in z] == [(B, m.shape[1], 192) for m in mp] and z[0].dtype == torch.bfloat16
        loss = vj.jepa_loss(z, h)
        assert loss.shape == () and loss.requires_grad
        vj.reg_loss(z)
        loss.backward()
        for net in (enc, pred):
            for n, p in net.named_parameters():
                if p.requires_grad:
                    assert p.grad is not None and p.grad.shape == p.shape, n
                else:
                    assert p.grad is None, n
        opt.step()
        opt.zero_grad()
        vj.ema_update(enc, tgt, 0.998)
    # every kernel family was reached
    for k in ("vj_gemm", "vj_attn_fwd", "vj_attn_bwd", "vj_layernorm_fwd", "vj_layernorm_bwd", "vj_im2col_tubelets",
              "vj_target_ln_gather", "vj_pred_assemble_fwd", "vj_pred_assemble_bwd", "vj_seq_slice", "vj_l1_loss_fwd",
              "vj_l1_loss_bwd", "vj_colsum", "vj_cast_f32_bf16", "vj_ema_update_shadow", "vj_adamw_step", "vj_token_std_accum"):
        assert k in dry, k
    # flat store survived deepcopy / re-adoption: parameters alias their store, target has its own buffer
    eb, tb = enc.backbone, tgt.backbone
    assert eb._store.flat.data_ptr() != tb._store.flat.data_ptr()
    assert all(p.data_ptr() >= eb._store.flat.data_ptr() for p in eb.parameters())
    # state_dict round trip keeps reference keys
    sd = enc.state_dict()
    assert "backbone.blocks.0.attn.qkv.weight" in sd and "backbone.pos_embed" in sd
    enc.load_state_dict(sd)
    osd = opt.state_dict()
    assert len(osd["param_groups"]) == 4 and "exp_avg" in next(iter(osd["state"].values()))


def test_padded_head_geometry(dry, monkeypatch):
    """ViT-L predictor: 384 / 16 heads = 24 -> padded to 32 (qkv N = 3*16*32 = 1536)."""
    _adopt_on_cpu(monkeypatch)
    from jepa_b200.models import vit_predictor
    p = vit_predictor(img_size=224, use_mask_tokens=True, patch_size=16, num_frames=16, tubelet_size=2, embed_dim=1024,
                      predictor_embed_dim=384, depth=1, num_heads=16, uniform_power=True, num_mask_tokens=2)
    assert p._spec.hd == 24 and p._spec.hdp == 32 and p._spec.padded and p._spec.inner == 512
    B, Ke, Kp = 2, 16, 24
    z = torch.zeros(B * Ke, 1024, dtype=torch.bfloat16, requires_grad=True)
    mc = torch.arange(Ke).repeat(B, 1)
    mt = torch.arange(Kp).repeat(B, 1) + Ke
    ctx = z.view(B, Ke, 1024)
    out = p.forward_multi([ctx], [mc], [mt], [0])
    assert tuple(out[0].shape) == (B, Kp, 1024)
    out[0].float().sum().backward()
    assert p.predictor_blocks[0].attn.qkv.weight.grad.shape == (1152, 384)
    assert dry.count("vj_head_pad") >= 6
