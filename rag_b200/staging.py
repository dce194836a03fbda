"""Eval-time input staging on the GPU: uint8 HWC image ->
ImageNet-normalised fp32 CHW, zero-padded on the top and right to the network's input size.

Reference: src/dataloaders/data_io.py:6-13 (ToTensor + Normalize) and
src/dataloaders/stereo_dataset.py:88-102 (``np.lib.pad`` to 480x960 -- removed in numpy 2.x)."""
from __future__ import annotations

import torch

from . import _cabi
from .functional import _stream


def normalize_pad(img_u8: torch.Tensor, out_h: int = 480, out_w: int = 960) -> torch.Tensor:
    """img_u8 [B,H,W,3] uint8 CUDA -> [B,3,out_h,out_w] fp32 (top_pad = out_h-H, right_pad = out_w-W)."""
    if not img_u8.is_cuda or img_u8.dtype != torch.uint8 or img_u8.dim() != 4 or img_u8.shape[-1] != 3:
        raise RuntimeError("rag_b200: normalize_pad wants a CUDA uint8 [B,H,W,3] tensor (no CPU fallback)")
    b, h, w, _ = img_u8.shape
    top, right = out_h - h, out_w - w
    if top < 0 or right < 0:
        raise RuntimeError(f"rag_b200: image {h}x{w} is larger than the padded size {out_h}x{out_w}")
    img_u8 = img_u8.contiguous()
    out = torch.empty((b, 3, out_h, out_w), dtype=torch.float32, device=img_u8.device)
    with torch.cuda.device(img_u8.device):
        rc = _cabi.lib().rag_normalize_pad(img_u8.data_ptr(), out.data_ptr(), b, h, w, top, right, _stream(img_u8))
    _cabi.check(rc, "rag_normalize_pad")
    return out
