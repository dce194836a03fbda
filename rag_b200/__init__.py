"""rag_b200 -- B200 (sm_100a) implementation of the stereo hot path of chzhang18/RAG:
concatenation cost volume + fused trilinear-upsample/softmin/soft-argmin disparity head,
forward and backward, behind the reference's nn.Module interfaces.  CUDA-only: no fallback."""
from .modules import CostVolume, Disp, DisparityRegression  # noqa: F401
from .functional import cost_volume, disp_head  # noqa: F401

__all__ = ["CostVolume", "Disp", "DisparityRegression", "cost_volume", "disp_head"]
