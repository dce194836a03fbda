"""Timing of the fused stem forward at the two BASELINE shapes (tuning tool)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_b200 import fused_stem as FS
ev = lambda: torch.cuda.Event(enable_timing=True)
g = torch.Generator(device="cuda").manual_seed(3)
for (b, hf, wf, md) in [(8, 160, 320, 192), (4, 96, 192, 192), (4, 128, 416, 288)]:
    x = torch.randn(b, 12, hf, wf, device="cuda", generator=g); y = torch.randn(b, 12, hf, wf, device="cuda", generator=g)
    w = torch.randn(12, 24, 3, 3, 3, device="cuda", generator=g) * 0.1
    sc = torch.rand(12, device="cuda", generator=g) + 0.5; sh = torch.randn(12, device="cuda", generator=g)
    f = lambda: FS.cv_stem_forward(x, y, w, sc, sh, True, md)
    ref = FS.cv_stem_forward(x, y, w, sc, sh, True, md, variant=1)
    out = f()
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for r in range(3):
        e0, e1 = ev(), ev(); e0.record()
        for _ in range(20): f()
        e1.record(); torch.cuda.synchronize(); ts.append(round(e0.elapsed_time(e1)/20, 4))
    print(json.dumps({"shape": [b, hf, wf], "ms": ts, "maxrel_vs_scalar_variant": float((out - ref).abs().max() / ref.abs().max())}))
