"""Pin the oracle: its fp32 C1 step must reproduce what the UNMODIFIED reference produced on the same
seeded weights / clips / masks (tests/golden/golden_step_c1.pt, written by tests/golden/make_golden.py)."""
import os
import sys

import pytest
import torch

from parity_util import run_c1_step_oracle, rel_l2

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "golden_step_c1.pt"))


@pytest.fixture(scope="module")
def oracle_step():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    return run_c1_step_oracle()


def test_losses(gold, oracle_step):
    assert abs(oracle_step["loss_jepa"] - gold["loss_jepa"]) < 2e-5
    assert abs(oracle_step["loss_reg"] - gold["loss_reg"]) < 2e-5


def test_activations(gold, oracle_step):
    for key, gkey in (("h", "h"), ("z", "z"), ("z_enc", "zenc")):
        for i, t in enumerate(oracle_step[key]):
            assert rel_l2(t[:, :4, :16], gold[f"{gkey}_slices"][i]) < 2e-4, key
            assert abs(float(t.norm()) - gold[f"{gkey}_norm"][i]) / gold[f"{gkey}_norm"][i] < 2e-4, key


def test_gradients(gold, oracle_step):
    for key in ("enc", "pred"):
        norms, slices = gold[f"{key}_grad_norm"], gold[f"{key}_grad_slices"]
        grads = oracle_step[f"{key}_grad"]
        assert set(norms) == set(grads)
        for n, ref in norms.items():
            assert abs(float(grads[n].norm()) - ref) <= 2e-3 * ref + 1e-9, (n, float(grads[n].norm()), ref)
        for n, ref in slices.items():
            got = grads[n].reshape(-1)[:32]
            assert float((got - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 1e-9, n


def test_ema(gold, oracle_step):
    for n, ref in gold["ema_slices"].items():
        assert torch.equal(oracle_step["ema"][n].reshape(-1)[:32], ref), n


def test_fused_mode_agrees_with_plain_restatement(oracle_step):
    """bench.py's CPU legs run the oracle with torch's fused CPU operators; same numbers as the plain path."""
    from oracle import vjepa_oracle as O
    O.FUSED = True
    try:
        fused = run_c1_step_oracle()
    finally:
        O.FUSED = False
    assert abs(fused["loss_jepa"] - oracle_step["loss_jepa"]) < 2e-5
    for key in ("h", "z"):
        for a, b in zip(fused[key], oracle_step[key]):
            assert rel_l2(a, b) < 1e-4
    for n, g in oracle_step["enc_grad"].items():
        assert rel_l2(fused["enc_grad"][n], g) < 2e-3, n


def test_input_pipeline_decisions_and_pixels_match_reference(golden_dir):
    """f3: the host sampler reproduces the reference's crop box / flip bit-exactly (same RNG call order), and the oracle's
    explicit bilinear restatement reproduces the reference transform's pixels (golden_transforms.pt was produced by the
    unmodified reference VideoTransform)."""
    import os
    import random
    import numpy as np
    from jepa_b200.transforms import make_transforms
    from oracle import vjepa_oracle as O
    sys_path_note = None  # noqa: F841
    cases = torch.load(os.path.join(golden_dir, "golden_transforms.pt"))
    for c in cases:
        T, H, W = c["shape"]
        buf = np.random.RandomState(1000 + c["seed"]).randint(0, 256, size=(T, H, W, 3), dtype=np.uint8)
        tf = make_transforms(random_horizontal_flip=True, random_resize_aspect_ratio=c["ratio"],
                             random_resize_scale=c["scale"], crop_size=c["crop"])
        random.seed(c["seed"]); np.random.seed(c["seed"])
        ticket = tf(buf)
        assert tuple(ticket.box) == tuple(c["box"]) and ticket.flip == c["flip"], (c["seed"], ticket.box, c["box"])
        assert torch.equal(ticket.frames, torch.from_numpy(buf))          # the frames travel untouched
        y = O.video_transform(buf, ticket.box, ticket.flip, c["crop"])
        assert y.shape == c["out"].shape
        assert float((y - c["out"]).abs().max()) < 2e-5, c["seed"]


# ------------------------------------------------------------------------------------------------ attentive probe (f4)
def _pooler_fixture():
    return torch.load(os.path.join(GOLDEN, "golden_pooler.pt"), weights_only=False)


def _build_probe(case):
    """jepa_b200.pooler.AttentiveClassifier built exactly like the fixture's reference module (same seed, same perturbation)."""
    from jepa_b200.pooler import AttentiveClassifier
    torch.manual_seed(case["seed"])
    clf = AttentiveClassifier(depth=1, **case["cfg"]).eval()
    with torch.no_grad():
        for n, p in clf.named_parameters():
            if n.endswith("bias") or "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
    return clf


def test_attentive_probe_state_dict_and_init_match_reference():
    """Same keys, same order, and - draw for draw - the same initial values as src/models/attentive_pooler.py."""
    sys.path.insert(0, os.path.join(GOLDEN))
    from common import sha16
    for case in _pooler_fixture()["cases"]:
        sd = _build_probe(case).state_dict()
        assert list(sd.keys()) == case["keys"]
        for k, v in sd.items():
            assert sha16(v) == case["sha"][k], k


def test_oracle_attentive_probe_matches_reference_fixture():
    from oracle import vjepa_oracle as O
    for case in _pooler_fixture()["cases"]:
        S = {k: v.double() for k, v in _build_probe(case).state_dict().items()}
        cfg = case["cfg"]
        pooled = O.attentive_pooler(S, case["x"].double(), cfg["num_heads"], cfg["complete_block"])
        logits = O.attentive_classifier(S, case["x"].double(), cfg["num_heads"], cfg["complete_block"])
        assert float((pooled.float() - case["pooled"]).abs().max()) < 2e-5
        assert float((logits.float() - case["logits"]).abs().max()) < 2e-5


def test_clip_aggregation_matches_reference_fixture():
    """evals/video_classification_frozen/utils.py:86-159 regrouping (pure tensor plumbing, runs on CPU tensors)."""
    from jepa_b200.pooler import ClipAggregation
    fx = _pooler_fixture()["clipagg"]

    class _Enc(torch.nn.Module):
        embed_dim, num_heads = 16, 2

        def forward(self, x):
            B, C, T, H, W = x.shape
            t = x.reshape(B, C, T // 2, 2, H // 4, 4, W // 4, 4).mean(dim=(3, 5, 7)).flatten(2).transpose(1, 2)
            return torch.cat([t * (i + 1) for i in range(6)], dim=-1)[..., :16]

    with torch.no_grad():
        out = ClipAggregation(_Enc(), tubelet_size=2, max_frames=64, use_pos_embed=True, attend_across_segments=True)(
            fx["xs1"], clip_indices=fx["idx"])
        out2 = ClipAggregation(_Enc(), tubelet_size=2, attend_across_segments=False)(fx["xs2"])
        out3 = ClipAggregation(_Enc(), tubelet_size=2, attend_across_segments=True)(fx["xs2"])
    assert all(torch.allclose(a, b, atol=1e-6) for a, b in zip(out, fx["out"]))
    assert all(torch.allclose(a, b, atol=1e-6) for va, vb in zip(out2, fx["out_noattend"]) for a, b in zip(va, vb))
    assert all(torch.allclose(a, b, atol=1e-6) for a, b in zip(out3, fx["out_attend_nopos"]))
