"""the Matching Net's last layer (Conv3d C -> 1, 3x3x3) on the hand-written kernel."""
import pytest
import torch
import torch.nn.functional as F

from tests._util import gen, randn

pytestmark = pytest.mark.gpu

SHAPES = [
    # (B, C, D, H, W)
    (1, 12, 64, 16, 64),      # the layer's real depth/channels, one w tile
    (2, 12, 20, 9, 68),       # ragged d / h / w tiles (W % 64 != 0, H % 8 != 0, D % 16 != 0)
    (1, 3, 5, 3, 4),          # smaller than one tile in every axis
    (1, 12, 17, 8, 132),      # two-and-a-bit w tiles, d tile boundary at 16
    (1, 24, 33, 11, 200),     # another channel count
]

# Shapes whose grid is deep enough for the march kernel (one warp per 8 x 8 x 64 column; conv3d_c1_fwd picks it from
# five warps per SM up): ragged h / w tiles, depth tails of one and two planes, one channel, the real channel count
MARCH_SHAPES = [
    (6, 3, 17, 100, 200),     # D % 8 == 1: the last column holds ONE output plane; H % 8 != 0, W % 64 != 0
    (8, 1, 26, 72, 256),      # a single channel (the copy ring runs ahead across planes); D % 8 == 2
    (4, 4, 56, 96, 192),      # the training crop's h x w
    (2, 12, 64, 96, 320),     # the layer's real depth and channels
    (111, 2, 8, 64, 64),      # many pairs, ONE column of planes each: both out-of-volume planes in every column
]


def _march_warps(shape):
    b, _, d, h, w = shape
    return ((w + 63) // 64) * ((h + 7) // 8) * b * ((d + 7) // 8)


@pytest.mark.parametrize("shape", SHAPES + MARCH_SHAPES, ids=str)
def test_conv3d_c1_matches_fp32_reference(shape):
    from rag_b200.last_conv import conv3d_c1_forward

    b, c, d, h, w = shape
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    assert (_march_warps(shape) >= 5 * sms) == (shape in MARCH_SHAPES), "the case no longer exercises the kernel it was written for"
    g = gen(sum(shape))
    x = randn((b, c, d, h, w), g)
    wt = randn((1, c, 3, 3, 3), g) * 0.1
    ref = F.conv3d(x.double(), wt.double(), padding=1)            # CPU fp64: the exact value of the reference's formula
    ref32 = F.conv3d(x, wt, padding=1)                            # CPU fp32: the reference's own arithmetic
    out = conv3d_c1_forward(x.cuda(), wt.cuda()).cpu()
    scale = ref.abs().max().item()
    assert (out.double() - ref).abs().max().item() <= 2e-6 * scale
    assert (out - ref32).abs().max().item() <= 4e-6 * scale
    # zero padding really is zero: an input that is non-zero only on the border planes
    xb = torch.zeros_like(x)
    xb[:, :, 0], xb[:, :, -1], xb[:, :, :, 0], xb[:, :, :, -1], xb[..., 0], xb[..., -1] = 1, 1, 1, 1, 1, 1
    outb = conv3d_c1_forward(xb.cuda(), wt.cuda()).cpu()
    assert (outb - F.conv3d(xb, wt, padding=1)).abs().max().item() <= 4e-6 * max(1.0, wt.abs().sum().item())


def test_convbr_dropin_and_fallbacks():
    """ConvBR_3d(C, 1, 3, 1, 1, bn=False, relu=False) through the patched forward: kernel under no_grad, the
    reference's nn.Conv3d whenever a gradient is wanted or the layer is not the 1-output-channel one."""
    import torch.nn as nn

    from rag_b200 import _cabi
    from rag_b200.fused_stem import stem_forward
    from rag_b200.last_conv import qualifies

    class ConvBR_3d(nn.Module):          # mirror of src/automl/operations_3d.py:31-47
        def __init__(self, cin, cout, k, s, p, bn=True, relu=True):
            super().__init__()
            self.relu, self.use_bn = relu, bn
            self.conv = nn.Conv3d(cin, cout, k, stride=s, padding=p, bias=False)
            self.bn = nn.BatchNorm3d(cout)

        forward = stem_forward

    g = gen(3)
    last = ConvBR_3d(12, 1, 3, 1, 1, bn=False, relu=False).cuda()
    x = randn((1, 12, 16, 8, 64), g).cuda()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        n0 = _cabi.launch_count()
        with torch.no_grad():
            y = last(x)
        assert _cabi.launch_count() == n0 + 1                       # the hand-written kernel ran
        ref = last.conv(x)
        assert (y - ref).abs().max().item() <= 4e-6 * ref.abs().max().item()
        n0 = _cabi.launch_count()
        yg = last(x.requires_grad_(True))                           # gradient wanted -> Conv3dC1Fn (kernels in both directions)
        assert _cabi.launch_count() == n0 + 1 and yg.requires_grad and "Conv3dC1Fn" in type(yg.grad_fn).__name__
        yg.sum().backward()
        assert _cabi.launch_count() == n0 + 4                       # + data gradient, weight partials, weight final
        assert last.conv.weight.grad is not None and x.grad is not None
        xr = x.detach().clone().requires_grad_(True)
        last.conv.weight.grad = None
        last.conv(xr).sum().backward()                              # the reference's own nn.Conv3d (fp32 cuDNN)
        assert (x.grad - xr.grad).abs().max().item() <= 1e-5 * xr.grad.abs().max().item()
        other = ConvBR_3d(12, 12, 3, 1, 1).cuda().eval()
        with torch.no_grad():
            assert not qualifies(other.conv, x)
            other(x.detach())
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_errors_are_loud():
    from rag_b200.last_conv import conv3d_c1_forward

    with pytest.raises(RuntimeError):
        conv3d_c1_forward(torch.zeros(1, 4, 4, 4, 6, device="cuda"), torch.zeros(1, 4, 3, 3, 3, device="cuda"))   # W % 4 != 0
    with pytest.raises(RuntimeError):
        conv3d_c1_forward(torch.zeros(1, 4, 4, 4, 8), torch.zeros(1, 4, 3, 3, 3))                                  # CPU tensors
    with pytest.raises(RuntimeError):
        conv3d_c1_forward(torch.zeros(1, 4, 4, 4, 8, device="cuda"), torch.zeros(2, 4, 3, 3, 3, device="cuda"))   # not 1 output channel


def test_golden_from_the_reference_layer():
    """tests/golden/last3_*.npz: the reference's own ConvBR_3d(12, 1, 3, 1, 1, bn=False, relu=False) run on CPU."""
    import glob
    import os

    import numpy as np

    from rag_b200.last_conv import conv3d_c1_forward
    from tests.conftest import GOLDEN

    paths = sorted(glob.glob(os.path.join(GOLDEN, "last3_*.npz")))
    assert paths
    for path in paths:
        z = np.load(path)
        out = conv3d_c1_forward(torch.from_numpy(z["feat"]).cuda(), torch.from_numpy(z["weight"]).cuda()).cpu().numpy()
        assert np.abs(out - z["out"]).max() <= 4e-6 * np.abs(z["out"]).max()
