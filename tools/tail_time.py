"""Timing of the last_3_3d kernels (forward, data gradient, weight gradient) at the training shape (tuning tool)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200 import _cabi  # noqa: E402

ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
L = _cabi.lib()
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731


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


g = torch.Generator(device="cuda").manual_seed(5)
for (b, c, d, h, w) in [(4, 12, 64, 96, 192), (8, 12, 64, 160, 320)]:
    x = torch.randn(b, c, d, h, w, device="cuda", generator=g)
    wt = torch.randn(1, c, 3, 3, 3, device="cuda", generator=g) * 0.1
    go = torch.randn(b, 1, d, h, w, device="cuda", generator=g)
    gin = torch.empty_like(x)
    gw = torch.empty_like(wt)
    ws = torch.empty(int(L.rag_conv3d_c1_bwd_workspace_bytes(c)) // 4, device="cuda")
    rec = {"shape": [b, c, d, h, w]}
    rec["bwd_data_ms"] = t(lambda: _cabi.check(L.rag_conv3d_c1_bwd(go.data_ptr(), None, wt.data_ptr(), gin.data_ptr(), None, None, b, c, d, h, w, st()), "bwd"))
    rec["bwd_weight_ms"] = t(lambda: _cabi.check(L.rag_conv3d_c1_bwd(go.data_ptr(), x.data_ptr(), None, None, gw.data_ptr(), ws.data_ptr(), b, c, d, h, w, st()), "bwd"))
    x64, w64 = x.double().requires_grad_(True), wt.double().requires_grad_(True)
    torch.nn.functional.conv3d(x64[:1], w64, padding=1).backward(go[:1].double())
    _cabi.check(L.rag_conv3d_c1_bwd(go[:1].contiguous().data_ptr(), x[:1].contiguous().data_ptr(), None, None, gw.data_ptr(), ws.data_ptr(), 1, c, d, h, w, st()), "bwd")
    rec["gw_maxrel_vs_fp64(b=1)"] = float((gw.double() - w64.grad).abs().max() / w64.grad.abs().max())
    print(json.dumps(rec), flush=True)
