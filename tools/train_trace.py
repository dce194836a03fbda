"""Per-kernel start/end times of the two-stream training step in steady state (tuning tool).

    python tools/train_trace.py [--schedule coresident] [--steps 12]

Events are recorded on both streams around every launch (they cost ~1-2 us each, so absolute step times are a little longer
than in bench.py); printed in ms relative to the first event of the traced window."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200 import functional as F_  # noqa: E402
from rag_b200 import pipeline as P_  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--schedule", default="coresident")
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--batch", type=int, default=4)
    a = ap.parse_args()
    hf, wf, df, md, c, b = 96, 192, 64, 192, 12, a.batch
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(b, c, hf, wf, device=dev, generator=g)
    y = torch.randn(b, c, hf, wf, device=dev, generator=g)
    cl = torch.randn(b, 1, df, hf, wf, device=dev, generator=g)
    gd = torch.randn(b, 3 * hf, 3 * wf, device=dev, generator=g) * (torch.rand(b, 3 * hf, 3 * wf, device=dev, generator=g) < 0.3)
    gc = torch.randn(b, 2 * c, df, hf, wf, device=dev, generator=g)
    vcf, vhf, vcb, vhb = P_.SCHEDULES[a.schedule]
    sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    disp, stats = F_.disp_head_forward(cl, md, True)
    rec = []
    n_warm = 20
    torch.cuda.synchronize()
    t0 = ev()
    for k in range(n_warm + a.steps):
        tr = k >= n_warm
        with torch.cuda.stream(sA):
            if k == n_warm: t0.record()
            e = [ev() for _ in range(3)]
            if tr: e[0].record()
            F_.cost_volume_forward(x, y, df, variant=vcf)
            if tr: e[1].record()
            F_.cost_volume_backward(gc, c, variant=vcb)
            if tr: e[2].record()
            if tr: rec.append(("cv", k - n_warm, e))
        with torch.cuda.stream(sB):
            e = [ev() for _ in range(3)]
            if tr: e[0].record()
            d2, s2 = F_.disp_head_forward(cl, md, True, variant=vhf)
            if tr: e[1].record()
            F_.disp_head_backward(cl, gd, d2, s2, md, variant=vhb)
            if tr: e[2].record()
            if tr: rec.append(("head", k - n_warm, e))
    torch.cuda.synchronize()
    for name, k, e in rec:
        t = [round(t0.elapsed_time(x_), 4) for x_ in e]
        print(json.dumps({"stream": name, "step": k, "fwd": [t[0], t[1]], "bwd": [t[1], t[2]], "fwd_ms": round(t[1] - t[0], 4), "bwd_ms": round(t[2] - t[1], 4)}))
    ends = [t0.elapsed_time(e[2]) for _, _, e in rec]
    print(json.dumps({"schedule": a.schedule, "ms_per_step_traced": round(max(ends) / a.steps, 4)}))


if __name__ == "__main__":
    main()
