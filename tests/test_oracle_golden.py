"""Pins the oracle (oracle/rag_oracle.py) to golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import rag_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CV = sorted(glob.glob(os.path.join(GOLDEN, "cv_*.npz")))
HEAD = sorted(glob.glob(os.path.join(GOLDEN, "head_*.npz")))


def test_golden_present():
    assert len(CV) >= 4 and len(HEAD) >= 5


@pytest.mark.parametrize("path", CV, ids=os.path.basename)
def test_cost_volume_oracle_bit_exact(path):
    z = np.load(path)
    md = int(z["maxdisp"])
    x, y = torch.from_numpy(z["x"]), torch.from_numpy(z["y"])
    cost = O.cost_volume_ref(x, y, md)
    assert torch.equal(cost, torch.from_numpy(z["cost"]))
    assert np.array_equal(O.cost_volume_closed(z["x"], z["y"], md), z["cost"])


@pytest.mark.parametrize("path", CV, ids=os.path.basename)
def test_cost_volume_grad_oracle_bit_exact(path):
    z = np.load(path)
    md = int(z["maxdisp"])
    c = z["x"].shape[1]
    gx, gy = O.cost_volume_grad_ref(torch.from_numpy(z["gcost"]), c, md)
    assert torch.equal(gx, torch.from_numpy(z["gx"])) and torch.equal(gy, torch.from_numpy(z["gy"]))
    # closed form with DESCENDING-d sequential fp32 sums is bit-equal; ascending is not
    gxc, gyc = O.cost_volume_grad_closed(z["gcost"], c)
    assert np.array_equal(gxc, z["gx"]) and np.array_equal(gyc, z["gy"])


def test_cost_volume_grad_order_matters():
    z = np.load(os.path.join(GOLDEN, "cv_b1_c12_h4_w70_md192.npz"))
    g = z["gcost"]
    c = 12
    asc = g[:, :c].copy()
    acc = np.zeros_like(z["gx"])
    for d in range(g.shape[2]):
        acc[..., d:] = acc[..., d:] + asc[:, :, d, :, d:]
    assert not np.array_equal(acc, z["gx"])  # ascending order differs -> the golden pins the order


@pytest.mark.parametrize("path", HEAD, ids=os.path.basename)
def test_head_oracle_matches_reference(path):
    z = np.load(path)
    md = int(z["maxdisp"])
    cost = torch.from_numpy(z["cost"])
    disp, gcost = O.disp_head_grad_ref(cost, torch.from_numpy(z["gdisp"]), md)
    # same torch ops in the same order; allow for ISA-dependent vectorised exp on another host
    np.testing.assert_allclose(disp.numpy(), z["disp"], rtol=0, atol=2e-5)
    scale = np.abs(z["gcost"]).max()
    assert np.abs(gcost.numpy() - z["gcost"]).max() <= 1e-5 * scale
    reg = O.disparity_regression_ref(torch.from_numpy(z["p"]), md)
    np.testing.assert_allclose(reg.numpy(), z["reg"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("path", HEAD, ids=os.path.basename)
def test_head_f64_restatement_within_tolerance(path):
    """Independent numpy fp64 restatement agrees with the reference within the stated
    tolerances: 1e-4 px on the disparity (sigma<=1), 1e-5 max-norm relative on the gradient."""
    z = np.load(path)
    md = int(z["maxdisp"])
    disp, g = O.disp_head_grad_f64(z["cost"][:, 0], z["gdisp"], md)
    # The reference's OWN fp32 rounding noise (sequential fp32 sum of p_k*k at magnitude ~maxdisp/2)
    # is ~8e-5 px at maxdisp=192/sigma=1 and grows with maxdisp and sigma.
    tol = 1e-4
    if "_md288_" in path:
        tol = 1.5e-4
    if "_s5" in path:
        tol = 6e-4
    assert np.abs(disp - z["disp"]).max() <= tol
    scale = np.abs(z["gcost"]).max()
    assert np.abs(g - z["gcost"][:, 0]).max() <= 2e-5 * scale


def test_upsample_tables_x3_structure():
    """Phase structure of the x3 tables."""
    for n in (4, 64, 96, 160, 416):
        i0, i1, l0, l1 = O.upsample_tables(n, 3 * n)
        dst = np.arange(3 * n)
        r = np.maximum(dst - 1, 0) // 3
        assert np.array_equal(i0, np.where(dst == 0, 0, r))
        assert np.array_equal(i1, np.minimum(i0 + 1, n - 1))
        assert np.all(np.abs(l1[dst % 3 == 1]) < 1e-4)
        assert np.all(np.abs(l1[dst % 3 == 2] - 1 / 3) < 1e-4)
        assert np.all(np.abs(l1[(dst % 3 == 0) & (dst > 0)] - 2 / 3) < 1e-4)


def test_metrics_oracle_matches_reference():
    z = np.load(os.path.join(GOLDEN, "metrics_b3_h12_w20.npz"))
    out = O.loss_and_metrics_ref(torch.from_numpy(z["est"]), torch.from_numpy(z["gt"]), int(z["maxdisp"]))
    for k in ("loss", "EPE", "D1", "Thres1", "Thres2", "Thres3"):
        assert abs(out[k] - float(z[k])) <= 1e-6 * max(1.0, abs(float(z[k]))), k


def test_stage_oracle_matches_reference():
    z = np.load(os.path.join(GOLDEN, "stage_h10_w14.npz"))
    out = O.normalize_pad_ref(z["img"], int(z["top_pad"]), int(z["right_pad"]))
    np.testing.assert_allclose(out, z["out"], rtol=0, atol=1e-6)


def _bn(d, training):
    mean = torch.from_numpy(d["bn_mean0" if training else "bn_mean"].copy())
    var = torch.from_numpy(d["bn_var0" if training else "bn_var"].copy())
    return dict(weight=torch.from_numpy(d["bn_weight"]).requires_grad_(training), bias=torch.from_numpy(d["bn_bias"]).requires_grad_(training),
                mean=mean, var=var, eps=float(d["bn_eps"]), momentum=float(d["bn_momentum"]) if training else 0.1)


def test_stem_and_last_conv_oracle_match_the_reference_layers():
    """stem3d0 on the reference-built volume (eval) and last_3_3d: the reference's own ConvBR_3d outputs."""
    d = np.load(os.path.join(GOLDEN, "stem_b1_c12_h5_w12_md24.npz"))
    out = O.stem_ref(torch.from_numpy(d["x"]), torch.from_numpy(d["y"]), torch.from_numpy(d["weight"]), _bn(d, False), int(d["maxdisp"]))
    np.testing.assert_allclose(out.numpy(), d["out"], rtol=1e-5, atol=1e-6)
    d = np.load(os.path.join(GOLDEN, "last3_b1_c12_d18_h9_w12.npz"))
    out = O.last_conv_ref(torch.from_numpy(d["feat"]), torch.from_numpy(d["weight"]))
    np.testing.assert_allclose(out.numpy(), d["out"], rtol=1e-5, atol=1e-6)


def test_training_mode_stem_oracle_matches_reference_forward_backward_and_running_stats():
    """The pin for the next round's fused training path: batch-statistics BatchNorm,
    gradients w.r.t. both feature maps, the conv weight and the BN affine parameters, running stats after the step."""
    d = np.load(os.path.join(GOLDEN, "trainstem_b2_c12_h5_w12_md24.npz"))
    x = torch.from_numpy(d["x"]).requires_grad_(True)
    y = torch.from_numpy(d["y"]).requires_grad_(True)
    w = torch.from_numpy(d["weight"]).requires_grad_(True)
    bn = _bn(d, True)
    out = O.stem_ref(x, y, w, bn, int(d["maxdisp"]), training=True)
    np.testing.assert_allclose(out.detach().numpy(), d["out"], rtol=1e-5, atol=1e-6)
    out.backward(torch.from_numpy(d["gout"]))
    for got, name in ((x.grad, "gx"), (y.grad, "gy"), (w.grad, "gweight"), (bn["weight"].grad, "gbn_weight"), (bn["bias"].grad, "gbn_bias")):
        ref = d[name]
        assert np.abs(got.numpy() - ref).max() <= 1e-5 * np.abs(ref).max(), name
    np.testing.assert_allclose(bn["mean"].numpy(), d["bn_mean1"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(bn["var"].numpy(), d["bn_var1"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("shape", [(1, 12, 5, 12, 24), (2, 3, 4, 9, 15), (1, 2, 3, 6, 30)], ids=str)
def test_volume_free_stem_algebra(shape):
    """The algebra the fused stem kernel rests on: tap sums over 18 row maps, and LF + RF
    where every tap is valid, equal the 3-D convolution of the materialised volume (fp64, incl. Df > Wf)."""
    b, c, hf, wf, md = shape
    g = torch.Generator().manual_seed(hf * wf)
    x = torch.randn(b, c, hf, wf, generator=g)
    y = torch.randn(b, c, hf, wf, generator=g)
    w = torch.randn(5, 2 * c, 3, 3, 3, generator=g)
    vol = O.cost_volume_ref(x, y, md)
    want = torch.nn.functional.conv3d(vol.double(), w.double(), None, 1, 1).numpy()
    for collapsed in (False, True):
        got = O.stem_volume_free_f64(x.numpy(), y.numpy(), w.numpy(), md, collapsed=collapsed)
        assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max()), f"collapsed={collapsed}"
    d = np.load(os.path.join(GOLDEN, "stem_b1_c12_h5_w12_md24.npz"))
    z = O.stem_volume_free_f64(d["x"], d["y"], d["weight"], int(d["maxdisp"]))
    sc = d["bn_weight"] / np.sqrt(d["bn_var"] + float(d["bn_eps"]))
    out = np.maximum(z * sc[None, :, None, None, None] + (d["bn_bias"] - d["bn_mean"] * sc)[None, :, None, None, None], 0)
    np.testing.assert_allclose(out, d["out"], rtol=1e-5, atol=2e-6)      # the reference's own (fp32) layer output


def test_stem_batch_moments_from_rows_match_the_training_golden():
    """Batch statistics of the stem convolution from the LF/RF rows (window sums, no volume): equal to the moments of
    z itself, and to the batch mean / variance the reference's BatchNorm3d used in train() (running-stat update)."""
    d = np.load(os.path.join(GOLDEN, "trainstem_b2_c12_h5_w12_md24.npz"))
    md = int(d["maxdisp"])
    s1, s2, n = O.stem_batch_moments_f64(d["x"], d["y"], d["weight"], md)
    z = O.stem_volume_free_f64(d["x"], d["y"], d["weight"], md)
    np.testing.assert_allclose(s1, z.sum(axis=(0, 2, 3, 4)), rtol=1e-12, atol=1e-10)
    np.testing.assert_allclose(s2, (z ** 2).sum(axis=(0, 2, 3, 4)), rtol=1e-12)
    mean = s1 / n
    var_unbiased = (s2 / n - mean ** 2) * n / (n - 1)
    mom = float(d["bn_momentum"])
    np.testing.assert_allclose((1 - mom) * d["bn_mean0"] + mom * mean, d["bn_mean1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose((1 - mom) * d["bn_var0"] + mom * var_unbiased, d["bn_var1"], rtol=1e-5, atol=1e-6)


def test_volume_free_stem_backward_matches_the_training_golden():
    """gx, gy, gweight of the training-mode stem from 2-D triangular / diagonal sums of the conv-output gradient
    (no volume, no volume gradient) against the reference's autograd through the materialised volume."""
    d = np.load(os.path.join(GOLDEN, "trainstem_b2_c12_h5_w12_md24.npz"))
    md = int(d["maxdisp"])
    z = torch.from_numpy(O.stem_volume_free_f64(d["x"], d["y"], d["weight"], md)).requires_grad_(True)
    out = torch.nn.functional.relu(torch.nn.functional.batch_norm(
        z, None, None, torch.from_numpy(d["bn_weight"]).double(), torch.from_numpy(d["bn_bias"]).double(), True, 0.1, float(d["bn_eps"])))
    out.backward(torch.from_numpy(d["gout"]).double())
    gx, gy, gw = O.stem_backward_volume_free_f64(d["x"], d["y"], d["weight"], z.grad.numpy(), md)
    for got, name in ((gx, "gx"), (gy, "gy"), (gw, "gweight")):
        ref = d[name]
        assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max(), (name, np.abs(got - ref).max() / np.abs(ref).max())


def test_training_step_of_the_stem_volume_free_end_to_end():
    """Forward (batch-statistics BatchNorm + ReLU) and every gradient of the stem from the volume-free pieces alone."""
    d = np.load(os.path.join(GOLDEN, "trainstem_b2_c12_h5_w12_md24.npz"))
    out, gx, gy, gw, gg, gb, _ = O.stem_train_volume_free_f64(d["x"], d["y"], d["weight"], d["bn_weight"], d["bn_bias"],
                                                              float(d["bn_eps"]), d["gout"], int(d["maxdisp"]))
    np.testing.assert_allclose(out, d["out"], rtol=1e-5, atol=2e-6)
    for got, name in ((gx, "gx"), (gy, "gy"), (gw, "gweight"), (gg, "gbn_weight"), (gb, "gbn_bias")):
        ref = d[name]
        assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max(), (name, np.abs(got - ref).max() / np.abs(ref).max())
