"""Pin the oracle: its fp32 C1 step must reproduce what the UNMODIFIED reference produced on the same
seeded weights / clips / masks (tests/golden/golden_step_c1.pt, written by tests/golden/make_golden.py)."""
import os

import pytest
import torch

from parity_util import run_c1_step_oracle, rel_l2


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
