import numpy as np
import torch


def gen(seed):
    return torch.Generator().manual_seed(seed)


def randn(shape, g, sigma=1.0):
    return sigma * torch.randn(*shape, generator=g)


def wide_grad(shape, g):
    """N(0,1) * exp(3 N(0,1)): wide dynamic range exposes the accumulation order."""
    return torch.randn(*shape, generator=g) * torch.exp(3 * torch.randn(*shape, generator=g))


def sparse_grad(shape, g, keep=0.3):
    return torch.randn(*shape, generator=g) * (torch.rand(*shape, generator=g) < keep).float()


def maxnorm_rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-9)  # floor: an all-zero gradient (Dl=1) compares as equal
