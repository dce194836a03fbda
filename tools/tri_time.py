"""Timing of the trilinear resize kernels, forward and backward separately (tuning tool)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200 import upsample as U  # noqa: E402

ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 4)


g = torch.Generator(device="cuda").manual_seed(2)
for (shape, size) in [((4, 12, 32, 48, 96), (64, 96, 192)), ((4, 12, 16, 24, 48), (32, 48, 96)), ((8, 12, 32, 80, 160), (64, 160, 320))]:
    x = torch.randn(*shape, device="cuda", generator=g)
    go = torch.randn(*shape[:2], *size, device="cuda", generator=g)
    rec = {"in": list(shape), "out": list(size)}
    rec["fwd_ms"] = t(lambda: U._resize("rag_trilinear_resize_fwd", x, tuple(shape[:2]) + tuple(size), tuple(shape[2:]), tuple(size), True))
    rec["bwd_ms"] = t(lambda: U._resize("rag_trilinear_resize_bwd", go, tuple(shape), tuple(shape[2:]), tuple(size), True))
    rec["aten_fwd_ms"] = t(lambda: F.interpolate(x, size=size, mode="trilinear", align_corners=True))
    xr = x.clone().requires_grad_(True)
    out = F.interpolate(xr, size=size, mode="trilinear", align_corners=True)
    rec["aten_bwd_ms"] = t(lambda: torch.autograd.grad(out, xr, go, retain_graph=True))
    ours = U.trilinear_resize(x, size, True)
    rec["fwd_maxrel"] = float((ours - out.detach()).abs().max() / out.detach().abs().max())
    xo = x.clone().requires_grad_(True)
    gi = torch.autograd.grad(U.trilinear_resize(xo, size, True), xo, go)[0]
    gr = torch.autograd.grad(out, xr, go, retain_graph=True)[0]
    rec["bwd_maxrel"] = float((gi - gr).abs().max() / gr.abs().max())
    nbytes = 4 * (x.numel() + go.numel())
    rec["hbm_floor_ms_each"] = round(nbytes / 7.0e12 * 1e3, 4)
    print(json.dumps(rec), flush=True)
