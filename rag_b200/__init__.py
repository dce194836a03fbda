"""rag_b200 -- B200 (sm_100a) implementation of the stereo hot path of chzhang18/RAG:
concatenation cost volume + fused trilinear-upsample/softmin/soft-argmin disparity head,
forward and backward, behind the reference's nn.Module interfaces.  CUDA-only: no fallback."""

__all__ = ["CostVolume", "Disp", "DisparityRegression", "cost_volume", "disp_head"]


def __getattr__(name):  # lazy: `python -m rag_b200.build` must not import torch or dlopen the library
    if name in ("CostVolume", "Disp", "DisparityRegression"):
        from . import modules

        return getattr(modules, name)
    if name in ("cost_volume", "disp_head"):
        from . import functional

        return getattr(functional, name)
    raise AttributeError(f"module 'rag_b200' has no attribute {name!r}")
