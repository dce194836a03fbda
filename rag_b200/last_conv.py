"""The Matching Net's last layer on the hand-written kernels, forward and backward.

The reference produces the disparity head's input with ``self.last_3_3d[i]`` = ConvBR_3d(C, 1, 3, 1, 1,
bn=False, relu=False) (rag_model.py:269, applied at :361-365): a bias-free Conv3d C -> 1, 3x3x3.  cuDNN spends
6.6 ms (TF32) / 32 ms (fp32) per 8 pairs at 480x960 on it; ``csrc/last_conv.cu`` does it in fp32 at the
FP32/HBM floor.  ``conv_forward`` is what ``stem_forward`` (fused_stem.py, bound onto ``ConvBR_3d.forward``
by ``rag_b200.network.install``) calls for a plain tensor: it takes the kernel when the layer and the call
qualify and otherwise runs the reference's own ``nn.Conv3d``.  With autograd the layer runs as ``Conv3dC1Fn``:
``csrc/last_conv_bwd.cu`` gives the data gradient (the C-times larger tensor written once) and the weight gradient
(deterministic two-stage reduction).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _cabi
from .functional import _require, _stream


def conv3d_c1_forward(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """out = conv3d(x, weight, padding=1) for weight [1,C,3,3,3]; x [B,C,D,H,W] CUDA fp32 -> [B,1,D,H,W]."""
    _require(x, "x"), _require(weight, "weight")
    if x.dim() != 5:
        raise RuntimeError(f"rag_b200: conv3d_c1 wants a [B,C,D,H,W] input, got {tuple(x.shape)}")
    b, c, d, h, w = x.shape
    if tuple(weight.shape) != (1, c, 3, 3, 3):
        raise RuntimeError(f"rag_b200: conv3d_c1 wants a [1,{c},3,3,3] weight, got {tuple(weight.shape)}")
    x, weight = x.contiguous(), weight.contiguous()
    out = torch.empty((b, 1, d, h, w), dtype=torch.float32, device=x.device)
    if out.numel() == 0:
        return out
    L = _cabi.lib()
    with torch.cuda.device(x.device):
        rc = L.rag_conv3d_c1_fwd(x.data_ptr(), weight.data_ptr(), out.data_ptr(), b, c, d, h, w, _stream(x))
    _cabi.check(rc, "rag_conv3d_c1_fwd")
    return out


class Conv3dC1Fn(torch.autograd.Function):
    """out = conv3d(x, weight[1,C,3,3,3], padding=1) with gradients w.r.t. x and weight from csrc/last_conv_bwd.cu."""

    @staticmethod
    def forward(ctx, x, weight):
        ctx.save_for_backward(x, weight)
        return conv3d_c1_forward(x, weight)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        x, g, weight = x.contiguous(), g.contiguous(), weight.contiguous()
        b, c, d, h, w = x.shape
        L = _cabi.lib()
        need_x, need_w = ctx.needs_input_grad
        gin = torch.empty_like(x) if need_x else None
        gw = torch.empty_like(weight) if need_w else None
        with torch.cuda.device(x.device):
            ws = torch.empty(int(L.rag_conv3d_c1_bwd_workspace_bytes(c)) // 4, dtype=torch.float32, device=x.device) if need_w else None
            rc = L.rag_conv3d_c1_bwd(g.data_ptr(), x.data_ptr(), weight.data_ptr(), gin.data_ptr() if need_x else None,
                                     gw.data_ptr() if need_w else None, ws.data_ptr() if need_w else None, b, c, d, h, w, _stream(x))
        _cabi.check(rc, "rag_conv3d_c1_bwd")
        return gin, gw


def _layer_ok(conv: nn.Module, x: torch.Tensor) -> bool:
    """The layer / call csrc/last_conv.cu implements (``conv3d_c1_fwd`` preconditions: C <= 64, W % 4 == 0, B*ceil(D/16) <=
    65535, 16-byte aligned tensors, not inside a CUDA-graph capture)."""
    if not (isinstance(conv, nn.Conv3d) and conv.out_channels == 1 and conv.bias is None and conv.kernel_size == (3, 3, 3)
            and conv.stride == (1, 1, 1) and conv.padding == (1, 1, 1) and conv.dilation == (1, 1, 1) and conv.groups == 1
            and conv.padding_mode == "zeros" and conv.weight.dtype == torch.float32 and conv.weight.is_cuda
            and conv.in_channels <= 64):
        return False
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 5
            and x.shape[1] == conv.in_channels and x.numel() > 0 and x.shape[-1] % 4 == 0
            and x.shape[2] * x.shape[3] * x.shape[4] < 2 ** 31 and x.shape[0] * ((x.shape[2] + 15) // 16) <= 65535
            and x.shape[0] * ((x.shape[2] + 15) // 16) * ((x.shape[3] + 3) // 4) * ((x.shape[4] + 31) // 32) < 2 ** 31):   # conv3d_c1_bwd
        return False
    if x.is_contiguous() and x.data_ptr() % 16:
        return False
    return not torch.cuda.is_current_stream_capturing()      # the constant-memory weight slots are event-ordered


def _wants_grad(conv: nn.Module, x: torch.Tensor) -> bool:
    return torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad)


def qualifies(conv: nn.Module, x: torch.Tensor) -> bool:
    """True when ``conv(x)`` is the layer the forward kernel implements and no gradient is wanted."""
    return _layer_ok(conv, x) and not _wants_grad(conv, x)


def conv_forward(conv: nn.Module, x: torch.Tensor) -> torch.Tensor:
    if not _layer_ok(conv, x):
        return conv(x)
    if _wants_grad(conv, x):
        return Conv3dC1Fn.apply(x, conv.weight)
    return conv3d_c1_forward(x, conv.weight)
