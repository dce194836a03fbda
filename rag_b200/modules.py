"""nn.Module surface of the hot path -- same constructor/forward signatures as the reference's
``Disp`` / ``DisparityRegression`` (src/models/rag_model.py:18-44) plus a ``CostVolume`` module for
the nine inline lines at rag_model.py:375-383.

All three are parameter-free, buffer-free (``state_dict()`` is empty, so the reference's
checkpoints at run.py:194 are unaffected), deepcopy-able and picklable (approaches/rag.py:225,
utils.py:64-70 deep-copy models every epoch): no ctypes handle is ever stored on a module.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as F_


class CostVolume(nn.Module):
    """cost[b,c,d,h,w] = x[b,c,h,w]*(w>=d); cost[b,C+c,d,h,w] = y[b,c,h,w-d]*(w>=d); d < int(maxdisp/stride)."""

    def __init__(self, maxdisp: int = 192, stride: int = 3):
        super().__init__()
        self.maxdisp = maxdisp
        self.stride = stride

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return F_.CostVolumeFn.apply(x, y, int(self.maxdisp / self.stride))

    def extra_repr(self) -> str:
        return f"maxdisp={self.maxdisp}, stride={self.stride}"


class DisparityRegression(nn.Module):
    """Same interface as rag_model.py:18-29: [B,maxdisp,H,W] probabilities -> [B,H,W]."""

    def __init__(self, maxdisp):
        super().__init__()
        self.maxdisp = maxdisp

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.is_contiguous() == True  # noqa: E712  (the reference's own guard, rag_model.py:24)
        return F_.DisparityRegressionFn.apply(x, self.maxdisp)


class Disp(nn.Module):
    """Same interface as rag_model.py:32-44: [B,1,Dl,Hl,Wl] matching cost -> [B,3Hl,3Wl] disparity.

    Keeps the reference's submodule names ``softmax`` and ``disparity`` (both stateless) so code
    that introspects them still works; the fused kernel does not call them."""

    def __init__(self, maxdisp=192):
        super().__init__()
        self.maxdisp = maxdisp
        self.softmax = nn.Softmin(dim=1)
        self.disparity = DisparityRegression(maxdisp=self.maxdisp)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F_.DispHeadFn.apply(x, self.maxdisp)
