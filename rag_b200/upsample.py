"""Trilinear resize on the hand-written kernels: the ``upsample_6`` / ``upsample_12`` steps of the
Matching Net's tail.

The reference builds them inside ``matching()`` / ``search_matching()`` on every call (rag_model.py:356-357, :675-676):

    upsample_6  = nn.Upsample(size=x.size()[2:],       mode='trilinear', align_corners=True)
    upsample_12 = nn.Upsample(size=[d//2, h//2, w//2], mode='trilinear', align_corners=True)

so the drop-in is a subclass of ``nn.Upsample`` that ``rag_b200.network.install(upsample=True)`` makes the reference's modules
see under the name ``nn.Upsample`` (a proxy of ``torch.nn`` bound as ``rag_model.nn``; everything else resolves to
``torch.nn`` unchanged).  5-D CUDA fp32 trilinear calls with an explicit ``size`` take ``csrc/trilinear.cu`` -- PyTorch's
fp32 index arithmetic, and a backward that GATHERS in a fixed order (bitwise repeatable; ATen scatters with atomicAdd) --
anything else runs ``nn.Upsample.forward`` as before.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _cabi
from .functional import _require, _stream


def _resize(fn_name: str, src: torch.Tensor, dst_shape, in_size, out_size, align_corners: bool) -> torch.Tensor:
    b, c = src.shape[:2]
    out = torch.empty(dst_shape, dtype=torch.float32, device=src.device)
    if out.numel() == 0:
        return out
    fn = getattr(_cabi.lib(), fn_name)
    with torch.cuda.device(src.device):
        rc = fn(src.data_ptr(), out.data_ptr(), b * c, *in_size, *out_size, int(bool(align_corners)), _stream(src))
    _cabi.check(rc, fn_name)
    return out


class TrilinearResizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size, align_corners):
        _require(x, "x")
        if x.dim() != 5:
            raise RuntimeError(f"rag_b200: trilinear resize wants a [B,C,D,H,W] tensor, got {tuple(x.shape)}")
        x = x.contiguous()
        size = tuple(int(s) for s in size)
        ctx.in_size, ctx.out_size, ctx.align = tuple(x.shape[2:]), size, bool(align_corners)
        ctx.bc = tuple(x.shape[:2])
        return _resize("rag_trilinear_resize_fwd", x, ctx.bc + size, ctx.in_size, size, align_corners)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        g = g.contiguous()
        gin = _resize("rag_trilinear_resize_bwd", g, ctx.bc + ctx.in_size, ctx.in_size, ctx.out_size, ctx.align)
        return gin, None, None


def trilinear_resize(x: torch.Tensor, size, align_corners: bool = True) -> torch.Tensor:
    """F.interpolate(x, size, mode='trilinear', align_corners=...) for CUDA fp32 [B,C,D,H,W]."""
    return TrilinearResizeFn.apply(x, size, align_corners)


class Upsample(nn.Upsample):
    """Same constructor / attributes / state (none) as ``nn.Upsample``; deepcopy- and pickle-safe."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if (self.mode == "trilinear" and self.size is not None and self.scale_factor is None and isinstance(x, torch.Tensor)
                and x.is_cuda and x.dtype == torch.float32 and x.dim() == 5 and x.numel() > 0):
            size = (self.size,) * 3 if isinstance(self.size, int) else tuple(self.size)
            if len(size) == 3 and max(x.shape[2:]) < 2 ** 14 and max(size) < 2 ** 14:
                return TrilinearResizeFn.apply(x, size, bool(self.align_corners))
        return super().forward(x)


class NNProxy:
    """What ``install(upsample=True)`` binds as ``rag_model.nn``: ``torch.nn`` with ``Upsample`` replaced."""

    Upsample = Upsample

    def __getattr__(self, name):
        return getattr(nn, name)
