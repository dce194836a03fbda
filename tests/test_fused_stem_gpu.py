"""Fused cost volume + first Matching-Net convolution against the reference
composition: cost volume (rag_model.py:375-383) -> Conv3d(2C->O,3,pad 1,bias=False) -> BatchNorm3d(eval)
-> ReLU (operations_3d.py:31-47), evaluated in full fp32 (TF32 off) on the materialised oracle volume.

Tolerance: the fused kernel sums the same 648 products per output in a different order (all fp32):
max-norm relative 1e-5 of the conv output."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import rag_oracle as O
from tests._util import gen, randn

pytestmark = pytest.mark.gpu


class ConvBR_3d(nn.Module):
    """Same fields / forward as the reference's ConvBR_3d (operations_3d.py:31-47)."""

    def __init__(self, c_in, c_out, bn=True, relu=True):
        super().__init__()
        self.relu = relu
        self.use_bn = bn
        self.conv = nn.Conv3d(c_in, c_out, 3, stride=1, padding=1, bias=False)
        self.bn = nn.BatchNorm3d(c_out)

    def forward(self, x):
        x = self.conv(x)
        if self.use_bn:
            x = self.bn(x)
        if self.relu:
            x = F.relu(x, inplace=True)
        return x


@pytest.fixture(autouse=True)
def _fp32_convs():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def ref_stem(x, y, layer, maxdisp):
    return ConvBR_3d.forward(layer, O.cost_volume_ref(x, y, maxdisp))


SHAPES = [
    # (B, C, Hf, Wf, maxdisp, O)
    (2, 12, 6, 40, 48, 12),     # Df = 16
    (1, 12, 5, 96, 192, 12),    # Df = 64 > band, Wf % 4 == 0
    (1, 12, 4, 38, 30, 12),     # Wf % 4 == 2 -> scalar store path
    (1, 12, 3, 12, 48, 12),     # Wf < Df: mostly masked
    (1, 12, 7, 416, 288, 5),    # config-5 width, O != C
    (1, 12, 4, 16, 9, 12),      # Df = 3: classes only, no interior d
    (1, 12, 3, 16, 6, 12),      # Df = 2 -> direct kernel
    (1, 5, 3, 20, 24, 4),       # C != 12 -> direct kernel
]


@pytest.mark.parametrize("shape", SHAPES, ids=str)
def test_fused_stem_matches_reference(shape):
    from rag_b200.fused_stem import cv_stem_forward

    b, c, hf, wf, md, o = shape
    g = gen(hash(shape) % 991)
    x, y = randn((b, c, hf, wf), g).cuda(), randn((b, c, hf, wf), g).cuda()
    torch.manual_seed(1)
    layer = ConvBR_3d(2 * c, o).cuda().eval()
    with torch.no_grad():
        layer.bn.running_mean.normal_(0, 0.5)
        layer.bn.running_var.uniform_(0.5, 2.0)
        layer.bn.weight.normal_(1, 0.2)
        layer.bn.bias.normal_(0, 0.3)
        conv_ref = layer.conv(O.cost_volume_ref(x, y, md))
        full_ref = ref_stem(x, y, layer, md)
        scale = layer.bn.weight * torch.rsqrt(layer.bn.running_var + layer.bn.eps)
        shift = layer.bn.bias - layer.bn.running_mean * scale
        for v in (None, 0) + ((1,) if (c == 12 and int(md / 3) >= 3 and wf >= 8) else ()) + ((2,) if (c == 12 and int(md / 3) >= 3 and wf >= 8 and wf % 4 == 0) else ()):
            conv = cv_stem_forward(x, y, layer.conv.weight, maxdisp=md, variant=v)
            assert conv.shape == conv_ref.shape
            err = (conv - conv_ref).abs().max().item() / conv_ref.abs().max().item()
            assert err <= 1e-5, f"variant {v}: conv rel err {err}"
            full = cv_stem_forward(x, y, layer.conv.weight, scale, shift, True, md, variant=v)
            err = (full - full_ref).abs().max().item() / max(full_ref.abs().max().item(), 1e-6)
            assert err <= 1e-5, f"variant {v}: conv+bn+relu rel err {err}"


def test_dropin_stem_forward_paths():
    """stem_forward bound onto a ConvBR_3d-shaped module: fused in eval/no-grad, conv-only fused with
    batch-statistics BN, materialised fallback under autograd -- all equal to the reference composition."""
    from rag_b200.fused_stem import VirtualCostVolume, stem_forward

    g = gen(3)
    md = 48
    x, y = randn((2, 12, 5, 40), g).cuda(), randn((2, 12, 5, 40), g).cuda()
    torch.manual_seed(2)
    layer = ConvBR_3d(24, 12).cuda()
    with torch.no_grad():
        layer.bn.running_mean.normal_(0, 0.5)
        layer.bn.running_var.uniform_(0.5, 2.0)
    # eval, no grad: fully fused
    layer.eval()
    with torch.no_grad():
        out = stem_forward(layer, VirtualCostVolume(x, y, md))
        ref = ref_stem(x, y, layer, md)
    assert (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    # a plain tensor goes through the reference code path unchanged
    with torch.no_grad():
        vol = O.cost_volume_ref(x, y, md)
        assert torch.equal(stem_forward(layer, vol.clone()), ref_stem(x, y, layer, md))
    # training-mode BN, no grad: fused conv + torch BN with batch statistics (running stats updated identically)
    l1, l2 = copy.deepcopy(layer).train(), copy.deepcopy(layer).train()
    with torch.no_grad():
        a = stem_forward(l1, VirtualCostVolume(x, y, md))
        b_ = ref_stem(x, y, l2, md)
    assert (a - b_).abs().max().item() <= 2e-5 * b_.abs().max().item()
    assert torch.allclose(l1.bn.running_mean, l2.bn.running_mean, rtol=1e-5, atol=1e-6)
    # autograd: falls back to the materialised volume, gradients flow to the features and the weights
    xr, yr = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    l3 = copy.deepcopy(layer).train()
    stem_forward(l3, VirtualCostVolume(xr, yr, md)).sum().backward()
    assert xr.grad is not None and yr.grad is not None and l3.conv.weight.grad is not None


def test_network_forward_with_fused_stem():
    """End to end through the drop-in forward with install(fuse_stem=True)-like wiring on a stand-in network."""
    from rag_b200 import network as N
    from rag_b200.fused_stem import stem_forward
    from rag_b200.modules import Disp

    class Net(nn.Module):
        def __init__(self, md):
            super().__init__()
            self.maxdisp = md
            self.disp = Disp(md)
            self.f = nn.Conv2d(3, 12, 3, stride=3, bias=False)
            self.stem3d0 = nn.ModuleList([ConvBR_3d(24, 12)])
            self.last = nn.Conv3d(12, 1, 3, padding=1, bias=False)

        def feature(self, img, task_arch=None, path=None):
            return self.f(img)

        def matching(self, cost, task_arch=None, path=None):
            return self.last(self.stem3d0[0](cost))

    md = 48
    torch.manual_seed(4)
    net = Net(md).cuda().eval()
    g = gen(8)
    left, right = randn((2, 3, 24, 96), g).cuda(), randn((2, 3, 24, 96), g).cuda()
    with torch.no_grad():
        ref = O.disp_head_ref(net.matching(O.cost_volume_ref(net.feature(left), net.feature(right), md)), md)
        import types

        ops3d = types.SimpleNamespace(ConvBR_3d=ConvBR_3d)
        try:
            N.install(operations_3d=ops3d, fuse_stem=True)
            assert ConvBR_3d.forward is stem_forward
            out = N.network_forward(net, left, right, 0)
            net.rag_b200_fuse_stem = False                      # per-instance opt-out: the materialised path
            out_plain = N.network_forward(net, left, right, 0)
            del net.rag_b200_fuse_stem
            N.install(fuse_stem=False)                          # a later plain install() switches the default off
            assert N._FUSE_STEM is False
        finally:
            N.uninstall()
        assert ConvBR_3d.forward is not stem_forward
    assert (out_plain - ref).abs().max().item() <= 1e-4
    assert (out - ref).abs().max().item() <= 2e-4


def test_golden_from_the_reference_layer():
    """tests/golden/stem_*.npz: the reference's Network.forward volume fed to its own ConvBR_3d(2C, C, 3, 1, 1) in
    eval mode (CPU, fp32)."""
    import glob
    import os

    import numpy as np

    from rag_b200.fused_stem import cv_stem_forward
    from tests.conftest import GOLDEN

    paths = sorted(glob.glob(os.path.join(GOLDEN, "stem_b*.npz")))
    assert paths
    for path in paths:
        z = np.load(path)
        t = lambda k: torch.from_numpy(np.asarray(z[k], dtype=np.float32)).cuda()  # noqa: E731
        scale = t("bn_weight") * torch.rsqrt(t("bn_var") + float(z["bn_eps"]))
        shift = t("bn_bias") - t("bn_mean") * scale
        out = cv_stem_forward(t("x"), t("y"), t("weight"), scale, shift, True, int(z["maxdisp"])).cpu().numpy()
        assert np.abs(out - z["out"]).max() <= 1e-5 * np.abs(z["out"]).max()


def _torch_moments(x, y, w, md):
    """fp64 moments of the fp32 (TF32 off) convolution of the materialised volume, per output channel."""
    from rag_b200.functional import cost_volume

    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        z = F.conv3d(cost_volume(x, y, md), w, None, 1, 1).double()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    return z.mean(dim=(0, 2, 3, 4)), z.var(dim=(0, 2, 3, 4), unbiased=False), z.numel() // z.shape[1]


@pytest.mark.parametrize("shape", [(2, 5, 12, 24), (1, 7, 40, 48), (3, 4, 36, 192), (2, 6, 416, 96)], ids=str)
def test_batch_stats_without_the_volume(shape):
    """First piece of the training path: the batch mean / variance a training-mode BatchNorm3d
    after the stem would use, from the collapsed row maps, against the moments of the real convolution output."""
    from rag_b200.fused_stem import cv_stem_batch_stats

    b, hf, wf, md = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(b, 12, hf, wf, generator=g).cuda()
    y = torch.randn(b, 12, hf, wf, generator=g).cuda()
    w = (torch.randn(12, 24, 3, 3, 3, generator=g) * 0.1).cuda()
    mean, var, n = cv_stem_batch_stats(x, y, w, md)
    rmean, rvar, rn = _torch_moments(x, y, w, md)
    assert n == rn
    std = rvar.sqrt()
    assert ((mean - rmean).abs() <= 1e-5 * std + 1e-7).all(), (mean - rmean).abs().max().item()
    assert ((var - rvar).abs() <= 1e-5 * rvar).all(), ((var - rvar).abs() / rvar).max().item()
    m2, v2, _ = cv_stem_batch_stats(x, y, w, md)
    assert torch.equal(mean, m2) and torch.equal(var, v2)          # fixed-order reduction: bitwise repeatable


def test_batch_stats_match_the_reference_running_stat_update():
    """tests/golden/trainstem_*.npz: the reference's ConvBR_3d in train() updated its running statistics from the batch
    moments (momentum 0.1, unbiased variance); ours must reproduce that update."""
    import os

    import numpy as np

    from rag_b200.fused_stem import cv_stem_batch_stats
    from tests.conftest import GOLDEN

    z = np.load(os.path.join(GOLDEN, "trainstem_b2_c12_h5_w12_md24.npz"))
    x, y, w = (torch.from_numpy(z[k]).cuda() for k in ("x", "y", "weight"))
    mean, var, n = cv_stem_batch_stats(x, y, w, int(z["maxdisp"]))
    mom = float(z["bn_momentum"])
    mean1 = (1 - mom) * torch.from_numpy(z["bn_mean0"]).double() + mom * mean.cpu()
    var1 = (1 - mom) * torch.from_numpy(z["bn_var0"]).double() + mom * var.cpu() * n / (n - 1)
    np.testing.assert_allclose(mean1.numpy(), z["bn_mean1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(var1.numpy(), z["bn_var1"], rtol=1e-5, atol=1e-6)
