"""CPU suite (-m "not gpu"): host logic against the reference-generated golden fixtures, C-ABI surface,
and the no-CPU-fallback contract.  Nothing here launches a kernel."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from common import C1, VITL16_MASKS, sha16

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(golden_dir):
    with open(os.path.join(golden_dir, "golden_host.json")) as f:
        return json.load(f)


def _ensure_built():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "jepa_b200", "libvjepa_b200.so")):
        g.build()


def test_c_abi_exports_every_declared_symbol():
    _ensure_built()
    from jepa_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "vjepa_b200.h")).read()
    declared = set(re.findall(r"\b(vj_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in vjepa_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().vj_version() == 100


def test_bad_arguments_are_rejected_without_a_gpu():
    _ensure_built()
    from jepa_b200 import _lib
    lib = _lib.load()
    # null operands -> negative return code and an error string, never a crash
    rc = lib.vj_gemm(None, 0, 0, None, 0, 0, None, 0, 0, 128, 64, 64, None, 1.0, 0, None, 0, 0, None, 0, None, 0, 1, 0, None)
    assert rc < 0 and b"null" in lib.vj_last_error_string()
    rc = lib.vj_attn_fwd(None, None, None, None, 1, 1, 1, 64, 1, 1.0, None)
    assert rc < 0


def test_pos_embed_matches_reference(host):
    from jepa_b200.pos_embs import get_3d_sincos_pos_embed
    for key, ref in host["pos_embed"].items():
        D, grid, depth = map(int, key.split("_"))
        e = get_3d_sincos_pos_embed(D, grid, depth, cls_token=False, uniform_power=True)
        assert e.shape == (depth * grid * grid, D)
        assert sha16(torch.from_numpy(e).float()) == ref["sha"]
        assert abs(float(e.sum()) - ref["sum"]) < 1e-6


def test_oracle_pos_embed_matches_reference(host):
    from oracle.vjepa_oracle import pos_embed_3d
    for key, ref in host["pos_embed"].items():
        D, grid, depth = map(int, key.split("_"))
        assert sha16(torch.from_numpy(pos_embed_3d(D, grid, depth)).float()) == ref["sha"]


@pytest.mark.parametrize("tag,crop,T,B", [("c2", 224, 16, 32), ("c4", 224, 16, 24), ("c5", 384, 16, 10), ("c1", 224, 8, 2)])
def test_multiblock3d_masks_bit_exact(host, golden_dir, tag, crop, T, B):
    import hashlib
    from src.masks.multiblock3d import MaskCollator
    torch.manual_seed(0)
    coll = MaskCollator(cfgs_mask=VITL16_MASKS, crop_size=crop, num_frames=T, patch_size=16, tubelet_size=2)
    gold_pt = torch.load(os.path.join(golden_dir, "golden_masks.pt"))
    for call, ref in enumerate(host["masks"][tag]):
        batch, me, mp = coll([torch.zeros(1) for _ in range(B)])
        assert batch.shape[0] == B
        assert [int(t.shape[1]) for t in me] == ref["Ke"] and [int(t.shape[1]) for t in mp] == ref["Kp"]
        h = hashlib.sha256()
        for t in me + mp:
            assert t.dtype == torch.int64
            h.update(t.numpy().tobytes())
        assert h.hexdigest()[:16] == ref["sha"]
        assert me[0][0][:20].tolist() == ref["first20"]
        if tag == "c1":
            for a, b in zip(me + mp, gold_pt[f"c1_call{call}"]["enc"] + gold_pt[f"c1_call{call}"]["pred"]):
                assert torch.equal(a, b)
        # structural properties: ascending unique keep-indices, context and targets disjoint
        for e, p in zip(me, mp):
            assert (e[:, 1:] > e[:, :-1]).all() and (p[:, 1:] > p[:, :-1]).all()
            for r in range(B):
                assert not set(e[r].tolist()) & set(p[r].tolist())


def test_mask_collator_step_advances_seed(host):
    from src.masks.multiblock3d import MaskCollator
    torch.manual_seed(0)
    coll = MaskCollator(cfgs_mask=VITL16_MASKS, crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)
    coll.step()  # resume path: train.py fast-forwards the counter
    torch.manual_seed(0)
    _, me, _ = coll([torch.zeros(1) for _ in range(4)])
    assert me[0].shape[1] != 0


def test_random_tube_masks_bit_exact(host):
    from src.masks.random_tube import MaskCollator
    np.random.seed(0)
    coll = MaskCollator(cfgs_mask=[dict(ratio=0.9)], crop_size=224, num_frames=16, patch_size=16, tubelet_size=2)
    _, me, mp = coll([torch.zeros(1) for _ in range(4)])
    ref = host["random_tube"]
    assert (int(me[0].shape[1]), int(mp[0].shape[1])) == (ref["Ke"], ref["Kp"])
    assert sha16(me[0]) == ref["sha_enc"] and sha16(mp[0]) == ref["sha_pred"]


def test_default_collator():
    from src.masks.default import DefaultCollator
    out, a, b = DefaultCollator()([torch.ones(2), torch.zeros(2)])
    assert out.shape == (2, 2) and a is None and b is None


def test_tensor_helpers(host):
    from src.utils.tensors import repeat_interleave_batch, trunc_normal_
    from oracle.vjepa_oracle import repeat_interleave_batch as o_rib
    assert repeat_interleave_batch(torch.arange(6), 2, 2).tolist() == host["repeat_interleave"]
    assert o_rib(torch.arange(6), 2, 2).tolist() == host["repeat_interleave"]
    assert torch.equal(repeat_interleave_batch(torch.arange(6), 2, 1), torch.arange(6))
    t = torch.empty(64, 32)
    torch.manual_seed(3)
    trunc_normal_(t, std=0.02)
    assert sha16(t) == host["trunc_normal"]["sha"]


def test_apply_masks_cpu_matches_oracle():
    from src.masks.utils import apply_masks
    from oracle.vjepa_oracle import apply_masks as o_apply
    x = torch.randn(3, 10, 8)
    m = [torch.tensor([[0, 3, 9], [1, 2, 3], [4, 5, 6]]), torch.tensor([[7], [8], [0]])]
    assert all(torch.equal(a, b) for a, b in zip(apply_masks(x, m, concat=False), o_apply(x, m, concat=False)))
    assert torch.equal(apply_masks(x, m[:1]), o_apply(x, m[:1]))
    assert apply_masks(x, [], concat=False) == []


def test_schedules_match_reference(host):
    from src.utils.schedulers import CosineWDSchedule, WarmupCosineSchedule

    class Opt:
        def __init__(self):
            self.param_groups = [dict(lr=0., weight_decay=0.), dict(lr=0., weight_decay=0., WD_exclude=True)]

    opt = Opt()
    sch = WarmupCosineSchedule(opt, warmup_steps=12, start_lr=0.0002, ref_lr=0.000625, final_lr=1e-6, T_max=100)
    wds = CosineWDSchedule(opt, ref_wd=0.04, final_wd=0.4, T_max=100)
    assert [sch.step() for _ in range(100)] == host["schedules"]["lr"]
    assert [wds.step() for _ in range(100)] == host["schedules"]["wd"]
    assert opt.param_groups[1]["weight_decay"] == host["schedules"]["wd_excluded"]
    assert opt.param_groups[0]["lr"] == host["schedules"]["lr"][-1] == opt.param_groups[1]["lr"]


def _build_c1():
    from app.vjepa.utils import init_video_model
    return init_video_model(device=torch.device("cpu"), patch_size=C1["patch_size"], num_frames=C1["num_frames"],
                            tubelet_size=C1["tubelet_size"], model_name=C1["model_name"], crop_size=C1["crop_size"],
                            pred_depth=C1["pred_depth"], pred_embed_dim=C1["pred_embed_dim"], uniform_power=True,
                            use_mask_tokens=True, num_mask_tokens=2, zero_init_mask_tokens=True, use_sdpa=True)


def test_state_dict_contract_matches_reference(host):
    import copy
    enc, pred = _build_c1()
    for net, tag in ((enc, "encoder"), (pred, "predictor")):
        sd = net.backbone.state_dict()
        assert {k: list(v.shape) for k, v in sd.items()} == host["state_shapes"][tag]
        assert [n for n, _ in net.backbone.named_parameters()] == host["param_order"][tag]
        assert all(k.startswith("backbone.") for k in net.state_dict())
    assert not enc.backbone.pos_embed.requires_grad and not pred.backbone.predictor_pos_embed.requires_grad
    # factory re-init undoes the per-layer rescale (std back to 0.02), biases zero, mask tokens zero
    assert abs(float(enc.backbone.blocks[11].mlp.fc2.weight.std()) - 0.02) < 2e-3
    assert float(pred.backbone.mask_tokens[0].abs().sum()) == 0.0
    assert pred.backbone.predictor_blocks[0].attn.num_heads == enc.backbone.num_heads == 3
    tgt = copy.deepcopy(enc)
    assert all(torch.equal(a, b) and a.data_ptr() != b.data_ptr() for a, b in zip(enc.parameters(), tgt.parameters()))
    # optimizer grouping (app/vjepa/utils.py:173-194)
    from app.vjepa.utils import init_opt
    opt, scaler, sch, wds = init_opt(enc, pred, iterations_per_epoch=10, start_lr=1e-4, ref_lr=1e-3, warmup=1,
                                     num_epochs=2, wd=0.04, final_wd=0.4)
    assert len(opt.param_groups) == 4 and scaler is None
    assert all(p.ndim > 1 for p in opt.param_groups[0]["params"]) and all(p.ndim == 1 for p in opt.param_groups[2]["params"])
    assert opt.param_groups[2]["WD_exclude"] and opt.param_groups[2]["weight_decay"] == 0
    sch.step(); wds.step()
    assert opt.param_groups[0]["weight_decay"] > 0 and opt.param_groups[3]["weight_decay"] == 0


def test_no_cpu_fallback():
    enc, pred = _build_c1()
    clips = torch.zeros(1, 3, C1["num_frames"], C1["crop_size"], C1["crop_size"])
    with pytest.raises(RuntimeError, match="CUDA"):
        enc(clips)
    with pytest.raises(RuntimeError, match="CUDA"):
        enc(clips, [torch.zeros(1, 8, dtype=torch.int64)])
    from jepa_b200 import kernels
    from jepa_b200._lib import VJError
    with pytest.raises(VJError):
        kernels.cast_f32_bf16(torch.zeros(8), torch.zeros(8, dtype=torch.bfloat16))


def test_product_path_never_imports_oracle():
    for top in ("jepa_b200", "src", "app"):
        for dp, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dp, f)).read()
                    assert "oracle" not in text.replace("# oracle", ""), f"{dp}/{f} references the oracle"


def test_bench_flop_accounting_matches_survey():
    """bench.py's roofline numerators are the SURVEY.md section 8d / BASELINE.md section 2 figures for the seeded masks."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("vj_bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    N224 = 8 * 14 * 14
    assert bench.flops_per_clip(1024, 24, N224, [360, 48], [824, 1144]) == 2_438_801_522_688      # C2/C3 ViT-L/16
    assert bench.flops_per_clip(1280, 32, N224, [360, 48], [824, 1144]) == 4_455_977_058_304      # C4 ViT-H/16
    me, mp = bench.seeded_masks(224, 16, 32, seed=0)   # the masks those figures are quoted on
    assert [m.shape[1] for m in me] == [360, 48] and [m.shape[1] for m in mp] == [824, 1144]
    g = bench.gemm_flops_per_clip(1024, 24, N224, [360, 48], [824, 1144])
    f = bench.flops_per_clip(1024, 24, N224, [360, 48], [824, 1144])
    assert abs((f - g) / f - 0.179) < 2e-3    # attention share 17.9 %
