"""torch.autograd Functions over the C ABI.  PyTorch is plumbing here (device memory, streams,
autograd graph); every byte of arithmetic happens in librag_b200.so.

Inputs must be CUDA fp32; anything else raises -- there is no CPU or eager fallback.
"""
from __future__ import annotations

import torch

from . import _cabi


def _require(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"rag_b200: {name} is on {t.device}; the hot path is CUDA-only (no CPU fallback)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"rag_b200: {name} has dtype {t.dtype}; the hot path computes in fp32 like the reference")
    return t


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


# ------------------------------------------------------------------------------------------------
# cost volume
# ------------------------------------------------------------------------------------------------
def cost_volume_forward(x: torch.Tensor, y: torch.Tensor, df: int, variant: int | None = None, workspace: bool = True) -> torch.Tensor:
    """``workspace=True`` (default) hands the library a fresh 16-byte device buffer for the work counter of the
    persistent kernel (rag_cost_volume_fwd_ws; the caching allocator makes it stream-ordered and private to the
    launch, graph-capture safe); ``workspace=False`` calls the stateless one-CTA-per-item entry point."""
    _require(x, "x"), _require(y, "y")
    if x.dim() != 4 or x.shape != y.shape:
        raise RuntimeError(f"rag_b200: cost volume wants x,y of identical [B,C,Hf,Wf] shape, got {tuple(x.shape)} and {tuple(y.shape)}")
    if x.device != y.device:
        raise RuntimeError("rag_b200: x and y are on different devices")
    x, y = x.contiguous(), y.contiguous()
    b, c, hf, wf = x.shape
    cost = torch.empty((b, 2 * c, df, hf, wf), dtype=torch.float32, device=x.device)
    if cost.numel() == 0:
        return cost
    L = _cabi.lib()
    with torch.cuda.device(x.device):
        ws = torch.empty(4, dtype=torch.int32, device=x.device) if workspace else None
        wp = ws.data_ptr() if ws is not None else None
        if variant is None:
            if ws is None:
                rc = L.rag_cost_volume_fwd(x.data_ptr(), y.data_ptr(), cost.data_ptr(), b, c, df, hf, wf, _stream(x))
            else:
                rc = L.rag_cost_volume_fwd_ws(x.data_ptr(), y.data_ptr(), cost.data_ptr(), b, c, df, hf, wf, wp, _stream(x))
        else:
            rc = L.rag_cost_volume_fwd_v(x.data_ptr(), y.data_ptr(), cost.data_ptr(), b, c, df, hf, wf, wp, variant, _stream(x))
    _cabi.check(rc, "rag_cost_volume_fwd")
    return cost


def cost_volume_backward(gcost: torch.Tensor, c: int, variant: int | None = None):
    _require(gcost, "gcost")
    gcost = gcost.contiguous()
    b, c2, df, hf, wf = gcost.shape
    if c2 != 2 * c:
        raise RuntimeError(f"rag_b200: gcost has {c2} channels, expected {2 * c}")
    gx = torch.empty((b, c, hf, wf), dtype=torch.float32, device=gcost.device)
    gy = torch.empty_like(gx)
    if gx.numel() == 0:
        return gx, gy
    if df == 0:
        return gx.zero_(), gy.zero_()
    L = _cabi.lib()
    with torch.cuda.device(gcost.device):
        if variant is None:
            rc = L.rag_cost_volume_bwd(gcost.data_ptr(), gx.data_ptr(), gy.data_ptr(), b, c, df, hf, wf, _stream(gcost))
        else:
            rc = L.rag_cost_volume_bwd_v(gcost.data_ptr(), gx.data_ptr(), gy.data_ptr(), b, c, df, hf, wf, variant, _stream(gcost))
    _cabi.check(rc, "rag_cost_volume_bwd")
    return gx, gy


class CostVolumeFn(torch.autograd.Function):
    """cost = concat-volume(x, y); saves only shapes."""

    @staticmethod
    def forward(ctx, x, y, df):
        ctx.c = x.shape[1]
        return cost_volume_forward(x, y, df)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gcost):
        gx, gy = cost_volume_backward(gcost, ctx.c)
        return gx, gy, None


def cost_volume(x: torch.Tensor, y: torch.Tensor, maxdisp: int = 192) -> torch.Tensor:
    """Drop-in for the inline loop at rag_model.py:375-383: Df = int(maxdisp/3) as there."""
    return CostVolumeFn.apply(x, y, int(maxdisp / 3))


# ------------------------------------------------------------------------------------------------
# disparity head
# ------------------------------------------------------------------------------------------------
def _head_shape(cost_lr: torch.Tensor):
    if cost_lr.dim() == 5:
        if cost_lr.shape[1] != 1:
            raise RuntimeError(f"rag_b200: Disp wants [B,1,Dl,Hl,Wl], got {tuple(cost_lr.shape)}")
        b, _, dl, hl, wl = cost_lr.shape
    elif cost_lr.dim() == 4:
        b, dl, hl, wl = cost_lr.shape
    else:
        raise RuntimeError(f"rag_b200: Disp wants a 5-D [B,1,Dl,Hl,Wl] tensor, got {tuple(cost_lr.shape)}")
    return b, dl, hl, wl


def disp_head_forward(cost_lr: torch.Tensor, maxdisp: int, want_stats: bool = True, variant: int | None = None,
                      out: torch.Tensor | None = None):
    """``out``: optional caller-owned contiguous [B,3Hl,3Wl] CUDA fp32 tensor the disparity is written into."""
    _require(cost_lr, "cost_lr")
    cost_lr = cost_lr.contiguous()
    b, dl, hl, wl = _head_shape(cost_lr)
    if out is not None:
        _require(out, "out")
        if tuple(out.shape) != (b, 3 * hl, 3 * wl) or not out.is_contiguous() or out.device != cost_lr.device:
            raise RuntimeError(f"rag_b200: out must be a contiguous [{b},{3 * hl},{3 * wl}] tensor on {cost_lr.device}")
    disp = out if out is not None else torch.empty((b, 3 * hl, 3 * wl), dtype=torch.float32, device=cost_lr.device)
    stats = torch.empty((b, 2, 3 * hl, 3 * wl), dtype=torch.float32, device=cost_lr.device) if want_stats else None
    if disp.numel() == 0:
        return disp, stats
    L = _cabi.lib()
    sp = stats.data_ptr() if stats is not None else None
    with torch.cuda.device(cost_lr.device):
        if variant is None:
            rc = L.rag_disp_head_fwd(cost_lr.data_ptr(), disp.data_ptr(), sp, b, dl, hl, wl, maxdisp, _stream(cost_lr))
        else:
            rc = L.rag_disp_head_fwd_v(cost_lr.data_ptr(), disp.data_ptr(), sp, b, dl, hl, wl, maxdisp, variant, _stream(cost_lr))
    _cabi.check(rc, "rag_disp_head_fwd")
    return disp, stats


def disp_head_backward(cost_lr, gdisp, disp, stats, maxdisp: int, variant: int | None = None):
    _require(cost_lr, "cost_lr"), _require(gdisp, "gdisp")
    cost_lr, gdisp = cost_lr.contiguous(), gdisp.contiguous()
    b, dl, hl, wl = _head_shape(cost_lr)
    gcost = torch.empty_like(cost_lr)
    if gcost.numel() == 0:
        return gcost
    L = _cabi.lib()
    # work buffer of the scratch + combine A/B variant only (the default adds both parts in place)
    scratch = torch.empty_like(cost_lr) if (maxdisp == 3 * dl and variant == 2) else None
    sp = scratch.data_ptr() if scratch is not None else None
    with torch.cuda.device(cost_lr.device):
        if variant is None:
            rc = L.rag_disp_head_bwd(cost_lr.data_ptr(), gdisp.data_ptr(), disp.data_ptr(), stats.data_ptr(),
                                     gcost.data_ptr(), sp, b, dl, hl, wl, maxdisp, _stream(cost_lr))
        else:
            rc = L.rag_disp_head_bwd_v(cost_lr.data_ptr(), gdisp.data_ptr(), disp.data_ptr(), stats.data_ptr(),
                                       gcost.data_ptr(), sp, b, dl, hl, wl, maxdisp, variant, _stream(cost_lr))
    _cabi.check(rc, "rag_disp_head_bwd")
    return gcost


class DispHeadFn(torch.autograd.Function):
    """disp = soft-argmin(softmin(upsample(cost_lr))); saves cost_lr, disp and 2 stats planes."""

    @staticmethod
    def forward(ctx, cost_lr, maxdisp):
        need = cost_lr.requires_grad
        disp, stats = disp_head_forward(cost_lr, maxdisp, want_stats=need)
        ctx.maxdisp = maxdisp
        if need:
            ctx.save_for_backward(cost_lr, disp, stats)
        return disp

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gdisp):
        cost_lr, disp, stats = ctx.saved_tensors
        return disp_head_backward(cost_lr, gdisp, disp, stats, ctx.maxdisp), None


def disp_head(cost_lr: torch.Tensor, maxdisp: int = 192) -> torch.Tensor:
    """Drop-in for Disp.forward (rag_model.py:39-44)."""
    return DispHeadFn.apply(cost_lr, maxdisp)


class DisparityRegressionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, maxdisp):
        _require(p, "x")
        if p.dim() != 4 or p.shape[1] != maxdisp:
            raise RuntimeError(f"rag_b200: DisparityRegression wants [B,{maxdisp},H,W], got {tuple(p.shape)}")
        assert p.is_contiguous()  # mirrors the reference's assert (rag_model.py:24)
        b, d, h, w = p.shape
        out = torch.empty((b, h, w), dtype=torch.float32, device=p.device)
        ctx.shape = (b, d, h, w)
        if out.numel():
            with torch.cuda.device(p.device):
                rc = _cabi.lib().rag_disparity_regression_fwd(p.data_ptr(), out.data_ptr(), b, d, h, w, _stream(p))
            _cabi.check(rc, "rag_disparity_regression_fwd")
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        b, d, h, w = ctx.shape
        gout = gout.contiguous()
        gp = torch.empty((b, d, h, w), dtype=torch.float32, device=gout.device)
        if gp.numel():
            with torch.cuda.device(gout.device):
                rc = _cabi.lib().rag_disparity_regression_bwd(gout.data_ptr(), gp.data_ptr(), b, d, h, w, _stream(gout))
            _cabi.check(rc, "rag_disparity_regression_bwd")
        return gp, None


def upsample_trilinear(cost_lr: torch.Tensor, maxdisp: int, fma_index: bool = True) -> torch.Tensor:
    """Validation helper: the F.interpolate at rag_model.py:40 alone -> [B,maxdisp,3Hl,3Wl]."""
    _require(cost_lr, "cost_lr")
    cost_lr = cost_lr.contiguous()
    b, dl, hl, wl = _head_shape(cost_lr)
    out = torch.empty((b, maxdisp, 3 * hl, 3 * wl), dtype=torch.float32, device=cost_lr.device)
    with torch.cuda.device(cost_lr.device):
        rc = _cabi.lib().rag_upsample_trilinear(cost_lr.data_ptr(), out.data_ptr(), b, dl, hl, wl, maxdisp, int(fma_index), _stream(cost_lr))
    _cabi.check(rc, "rag_upsample_trilinear")
    return out
