"""Does coupling the two streams of the training step once per step pin a good phase? (tuning tool)

    python tools/couple_probe.py

OverlappedTrainPath with phase_lock off / on, six timed windows of 100 steps each, twice."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200 import pipeline as P_  # noqa: E402

hf, wf, df, md, c, b = 96, 192, 64, 192, 12, 4
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(b, c, hf, wf, device="cuda", generator=g)
y = torch.randn(b, c, hf, wf, device="cuda", generator=g)
cl = torch.randn(b, 1, df, hf, wf, device="cuda", generator=g)
gd = torch.randn(b, 3 * hf, 3 * wf, device="cuda", generator=g) * (torch.rand(b, 3 * hf, 3 * wf, device="cuda", generator=g) < 0.3)
gc = torch.randn(b, 2 * c, df, hf, wf, device="cuda", generator=g)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
for rep in range(2):
    for lock in (False, True):
        tp = P_.OverlappedTrainPath(md, "cuda")
        tp.phase_lock = lock
        res = []
        for r in range(6):
            for _ in range(10):
                o = tp.step(x, y, cl, gc, gd)
            tp.join()
            torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            for _ in range(100):
                o = tp.step(x, y, cl, gc, gd)
            tp.join()
            e1.record()
            torch.cuda.synchronize()
            res.append(round(e0.elapsed_time(e1) / 100, 4))
        print(json.dumps({"phase_lock": lock, "ms_per_step": res}), flush=True)
