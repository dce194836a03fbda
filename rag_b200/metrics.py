"""Masked loss + metrics on the disparity map, fused.

Reference: src/approaches/rag.py:210-211,418-430 and src/utilstool/metrics.py:22-65.  The reference
computes, per batch, a masked smooth-L1 over the whole batch plus five per-image metrics, with about
ten tiny kernels per image and a ``.item()`` sync per scalar.  Here one kernel pair produces eight
per-image sums; everything else is arithmetic on a [B,8] tensor.

sums[b] = (n_mask, n_pos, sum smooth_l1, sum |E|, n_D1, n_E>1, n_E>2, n_E>3), mask = 0 < gt < maxdisp.
"""
from __future__ import annotations

import torch

from . import _cabi
from .functional import _require, _stream

KEYS = ("loss", "EPE", "D1", "Thres1", "Thres2", "Thres3")


def loss_metric_sums(est: torch.Tensor, gt: torch.Tensor, maxdisp: float = 192.0) -> torch.Tensor:
    """est, gt [B,H,W] CUDA fp32 -> per-image sums [B,8] float64 on the same device."""
    _require(est, "disp_est"), _require(gt, "disp_gt")
    if est.dim() != 3 or est.shape != gt.shape:
        raise RuntimeError(f"rag_b200: loss/metrics want two [B,H,W] tensors, got {tuple(est.shape)} and {tuple(gt.shape)}")
    est, gt = est.contiguous(), gt.contiguous()
    b, h, w = est.shape
    L = _cabi.lib()
    sums = torch.empty((b, 8), dtype=torch.float64, device=est.device)
    scratch = torch.empty((b, L.rag_loss_metrics_scratch(h, w)), dtype=torch.float64, device=est.device)
    with torch.cuda.device(est.device):
        rc = L.rag_loss_metrics_sums(est.data_ptr(), gt.data_ptr(), sums.data_ptr(), scratch.data_ptr(), b, h, w, float(maxdisp), _stream(est))
    _cabi.check(rc, "rag_loss_metrics_sums")
    return sums


class _MaskedSmoothL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, est, gt, maxdisp):
        sums = loss_metric_sums(est.detach(), gt, maxdisp)
        ctx.save_for_backward(est.detach(), gt, sums)
        ctx.maxdisp = float(maxdisp)
        ctx.mark_non_differentiable(sums)
        loss = (sums[:, 2].sum() / sums[:, 0].sum()).to(torch.float32)
        return loss, sums

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gloss, _gsums):
        est, gt, sums = ctx.saved_tensors
        est, gt = est.contiguous(), gt.contiguous()
        b, h, w = est.shape
        gest = torch.empty_like(est)
        gl = gloss.to(torch.float32).contiguous()
        with torch.cuda.device(est.device):
            rc = _cabi.lib().rag_smooth_l1_bwd(est.data_ptr(), gt.data_ptr(), sums.data_ptr(), gl.data_ptr(), gest.data_ptr(),
                                               b, h, w, ctx.maxdisp, _stream(est))
        _cabi.check(rc, "rag_smooth_l1_bwd")
        return gest, None, None


def masked_smooth_l1(est: torch.Tensor, gt: torch.Tensor, maxdisp: float = 192.0):
    """Drop-in for ``F.smooth_l1_loss(est[mask], gt[mask], size_average=True)`` with
    ``mask = (gt < maxdisp) & (gt > 0)`` (approaches/rag.py:210-211).  Returns (loss, sums[B,8]);
    differentiable w.r.t. est."""
    return _MaskedSmoothL1.apply(est, gt, maxdisp)


def metrics_from_sums(sums: torch.Tensor) -> torch.Tensor:
    """[B,8] sums -> 6-vector (loss, EPE, D1, Thres1, Thres2, Thres3) with the reference's per-batch
    semantics: loss = masked mean over the whole batch; each metric = mean over the images that are
    NOT skipped (mask.mean()/(gt>0).mean() < 0.1 skips an image, metrics.py:31) of the per-image masked
    mean; if every image is skipped the metric is 0 (metrics.py:36-38).  Stays on the device."""
    n_mask, n_pos = sums[:, 0], sums[:, 1]
    keep = ~((n_mask / n_pos) < 0.1)          # NaN (n_pos == 0) compares False -> image kept, like the reference
    nk = keep.sum()
    per_img = sums[:, 3:8] / n_mask[:, None]  # EPE, D1, Thres1..3 per image
    per_img = torch.where(keep[:, None], per_img, torch.zeros_like(per_img))
    mets = torch.where(nk > 0, per_img.sum(0) / nk.clamp(min=1), torch.zeros(5, dtype=sums.dtype, device=sums.device))
    loss = sums[:, 2].sum() / n_mask.sum()
    return torch.cat([loss[None], mets])


class MetricAccumulator:
    """Replacement for the reference's ``AverageMeterDict`` over ``tensor2float`` scalars
    (utilstool/experiment.py:126-151): unweighted mean over batches of the per-batch values.
    Accumulates on the device (no ``.item()`` per scalar); ``mean()`` does one all-reduce of a
    7-element fp64 vector when a process group is given, then one device->host copy."""

    def __init__(self, device):
        self.acc = torch.zeros(7, dtype=torch.float64, device=device)  # 6 sums + batch count

    def update(self, sums: torch.Tensor) -> None:
        self.acc[:6] += metrics_from_sums(sums)
        self.acc[6] += 1

    def mean(self, group=None, distributed: bool | None = None) -> dict:
        import torch.distributed as dist

        acc = self.acc.clone()
        if distributed is None:
            distributed = dist.is_available() and dist.is_initialized()
        if distributed:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
        host = acc.cpu()
        n = float(host[6])
        return {k: (float(host[i]) / n if n > 0 else float("nan")) for i, k in enumerate(KEYS)}
