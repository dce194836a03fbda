"""The Matching Net's tail on the hand-written kernels: `upsample_12` / `upsample_6`
(nn.Upsample trilinear, align_corners=True; rag_model.py:356-366) forward + deterministic backward, and the backward of
`last_3_3d` (Conv3d C -> 1; rag_model.py:269).  Against the goldens captured from the reference's own `matching()` tail
(tests/golden/make_golden_tail.py), against PyTorch on this GPU, and against fp64."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests._util import gen, randn

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _mx(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def test_tail_goldens_from_the_reference():
    from rag_b200.last_conv import Conv3dC1Fn, conv3d_c1_forward
    from rag_b200.upsample import trilinear_resize

    z = np.load(os.path.join(GOLDEN, "tail_c12_d16_h12_w24.npz"))
    t = lambda k: torch.from_numpy(z[k]).cuda()  # noqa: E731
    for name in ("up12", "up6"):
        x = t(f"{name}_in").requires_grad_(True)
        out = trilinear_resize(x, t(f"{name}_out").shape[2:], align_corners=True)
        assert _mx(out.detach(), t(f"{name}_out")) <= 2e-6, name
        out.backward(t(f"{name}_gout"))
        assert _mx(x.grad, t(f"{name}_gin")) <= 2e-6, name
    x, w = t("conv_in").requires_grad_(True), t("conv_weight").requires_grad_(True)
    out = Conv3dC1Fn.apply(x, w)
    assert torch.equal(out.detach(), conv3d_c1_forward(x.detach(), w.detach()))
    assert _mx(out.detach(), t("conv_out")) <= 2e-6
    out.backward(t("conv_gout"))
    assert _mx(x.grad, t("conv_gin")) <= 2e-6
    assert _mx(w.grad, t("conv_gweight")) <= 1e-5


RESIZE_CASES = [
    # (B, C, in (D,H,W), out (D,H,W), align_corners)
    (2, 3, (4, 3, 6), (8, 6, 12), True),           # upsample_12 geometry
    (1, 12, (32, 48, 96), (64, 96, 192), True),    # upsample_6 at BASELINE config 3 (one pair)
    (1, 2, (5, 7, 9), (11, 13, 18), True),         # no simple ratio, Wo % 4 != 0
    (1, 2, (8, 6, 10), (4, 3, 5), True),           # DOWN-sampling (scale > 1: gaps in the source index table)
    (1, 1, (1, 4, 4), (1, 9, 7), True),            # singleton axis (scale 0 with align_corners)
    (2, 2, (6, 5, 8), (18, 15, 24), False),        # align_corners=False, x3 (the head's convention)
    (1, 1, (3, 3, 3), (3, 3, 3), True),            # identity
    (1, 1, (3, 4, 132), (5, 7, 264), True),        # row kernels, wide rows: two input vectors per lane, loads one row ahead
    (1, 2, (4, 4, 16), (8, 8, 32), False),         # row kernels with align_corners=False (x2)
    (1, 1, (2, 3, 8), (3, 5, 12), True),           # row kernels, ragged: 3 output vectors, one lane in four busy
]


@pytest.mark.parametrize("case", RESIZE_CASES, ids=str)
def test_trilinear_resize_vs_torch(case):
    from rag_b200.upsample import trilinear_resize

    b, c, si, so, ac = case
    g = gen(hash(case) % 991)
    x0 = randn((b, c) + si, g).cuda()
    go = randn((b, c) + so, g).cuda()
    xr = x0.clone().requires_grad_(True)
    ref = F.interpolate(xr, so, mode="trilinear", align_corners=ac)
    ref.backward(go)
    x64 = x0.double().requires_grad_(True)
    r64 = F.interpolate(x64, so, mode="trilinear", align_corners=ac)        # fp64 lambdas too: bounds both fp32 results
    r64.backward(go.double())
    x = x0.clone().requires_grad_(True)
    out = trilinear_resize(x, so, ac)
    out.backward(go)
    assert _mx(out.detach(), ref.detach()) <= 1e-6, f"fwd vs torch fp32: {_mx(out.detach(), ref.detach()):.2e}"
    assert _mx(x.grad, xr.grad) <= 2e-6, f"bwd vs torch fp32: {_mx(x.grad, xr.grad):.2e}"
    # fp64 uses exact lambdas; PyTorch's fp32 lambdas (reproduced here) are off by ~6e-8 * index, i.e. up to ~1e-5 at index 191
    assert _mx(out.detach(), r64.detach()) <= 2e-5 and _mx(x.grad, x64.grad) <= 2e-5
    # deterministic gather: bitwise repeatable
    x2 = x0.clone().requires_grad_(True)
    trilinear_resize(x2, so, ac).backward(go)
    assert torch.equal(x2.grad, x.grad)


def test_upsample_module_is_a_drop_in():
    import copy
    import pickle

    from rag_b200.upsample import NNProxy, Upsample

    nnp = NNProxy()
    assert nnp.Conv3d is torch.nn.Conv3d and nnp.Upsample is Upsample
    up = nnp.Upsample(size=[8, 6, 12], mode="trilinear", align_corners=True)
    assert isinstance(up, torch.nn.Upsample) and len(up.state_dict()) == 0
    up = pickle.loads(pickle.dumps(copy.deepcopy(up)))
    x = randn((1, 2, 4, 3, 6), gen(3)).cuda()
    assert _mx(up(x), F.interpolate(x, [8, 6, 12], mode="trilinear", align_corners=True)) <= 1e-6
    # anything the kernel does not cover runs nn.Upsample itself: 4-D bilinear, scale_factor, CPU tensors
    y = randn((1, 2, 5, 7), gen(4)).cuda()
    assert torch.equal(Upsample(size=[10, 14], mode="bilinear", align_corners=True)(y),
                       torch.nn.Upsample(size=[10, 14], mode="bilinear", align_corners=True)(y))
    assert torch.equal(Upsample(scale_factor=2, mode="trilinear", align_corners=True)(x),
                       torch.nn.Upsample(scale_factor=2, mode="trilinear", align_corners=True)(x))
    assert Upsample(size=[8, 6, 12], mode="trilinear", align_corners=True)(x.cpu()).device.type == "cpu"


CONV_CASES = [(1, 12, 16, 12, 24), (2, 12, 9, 5, 36), (1, 5, 3, 4, 8), (2, 12, 64, 24, 192), (1, 64, 6, 9, 40),
              (1, 3, 17, 7, 68),     # depth tail of ONE plane behind a full 16-plane segment (data gradient), ragged h / w
              (1, 2, 33, 5, 36)]     # the same behind a 32-plane segment (weight gradient)


@pytest.mark.parametrize("case", CONV_CASES, ids=str)
def test_last_conv_backward_vs_fp64(case):
    from rag_b200.last_conv import Conv3dC1Fn

    b, c, d, h, w = case
    g = gen(hash(case) % 983)
    x0 = randn((b, c, d, h, w), g).cuda()
    w0 = (randn((1, c, 3, 3, 3), g) * 0.2).cuda()
    go = randn((b, 1, d, h, w), g).cuda()
    x64, w64 = x0.double().requires_grad_(True), w0.double().requires_grad_(True)
    F.conv3d(x64, w64, padding=1).backward(go.double())
    x, wt = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    out = Conv3dC1Fn.apply(x, wt)
    out.backward(go)
    assert _mx(x.grad, x64.grad) <= 2e-6, f"data gradient {_mx(x.grad, x64.grad):.2e}"
    assert _mx(wt.grad, w64.grad) <= 1e-5, f"weight gradient {_mx(wt.grad, w64.grad):.2e}"
    # partial gradients + determinism
    x2 = x0.clone().requires_grad_(True)
    Conv3dC1Fn.apply(x2, w0).backward(go)
    assert torch.equal(x2.grad, x.grad)
    w2 = w0.clone().requires_grad_(True)
    Conv3dC1Fn.apply(x0, w2).backward(go)
    assert torch.equal(w2.grad, wt.grad)


def test_conv_forward_routes_training_through_the_kernels():
    from rag_b200 import _cabi
    from rag_b200.last_conv import conv_forward

    conv = torch.nn.Conv3d(12, 1, 3, padding=1, bias=False).cuda()
    x = randn((1, 12, 8, 6, 16), gen(9)).cuda().requires_grad_(True)
    n0 = _cabi.launch_count()
    out = conv_forward(conv, x)
    out.sum().backward()
    assert _cabi.launch_count() - n0 == 4          # forward, data gradient, weight partials, weight final
    ref = F.conv3d(x.detach().double(), conv.weight.detach().double(), padding=1)       # cuDNN's fp32 default is TF32: compare in fp64
    assert _mx(out.detach(), ref) <= 2e-6 and x.grad is not None and conv.weight.grad is not None
