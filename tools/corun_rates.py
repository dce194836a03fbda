"""How much two kernels slow each other when they share the GPU on two streams in STEADY STATE (tuning tool).

    python tools/corun_rates.py [--schedule coresident]

For each pair (volume kernel, head kernel): stream A runs the volume kernel N times back to back, stream B the head kernel
M times back to back, with N and M chosen so that both take about the same time alone; durations of each stream alone and of
each stream when both run together -> slow-down factors and the "sharing efficiency" (alone_A + alone_B) / together."""
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
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--cfg", default="288x576")
    a = ap.parse_args()
    hf, wf = {"288x576": (96, 192), "480x960": (160, 320)}[a.cfg]
    df, md, c, b = 64, 192, 12, a.batch
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(b, c, hf, wf, device=dev, generator=g)
    y = torch.randn(b, c, hf, wf, device=dev, generator=g)
    cl = torch.randn(b, 1, df, hf, wf, device=dev, generator=g)
    gd = torch.randn(b, 3 * hf, 3 * wf, device=dev, generator=g) * (torch.rand(b, 3 * hf, 3 * wf, device=dev, generator=g) < 0.3)
    gc = torch.randn(b, 2 * c, df, hf, wf, device=dev, generator=g)
    vcf, vhf, vcb, vhb = P_.SCHEDULES[a.schedule]
    disp, stats = F_.disp_head_forward(cl, md, True)
    kern = {
        "cv_fwd": lambda: F_.cost_volume_forward(x, y, df, variant=vcf),
        "cv_bwd": lambda: F_.cost_volume_backward(gc, c, variant=vcb),
        "head_fwd": lambda: F_.disp_head_forward(cl, md, True, variant=vhf),
        "head_bwd": lambda: F_.disp_head_backward(cl, gd, disp, stats, md, variant=vhb),
    }
    sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def run(fa, na, fb, nb):
        """returns (ms of stream A's na launches, ms of stream B's nb launches); fa or fb may be None"""
        torch.cuda.synchronize()
        a0, a1, b0, b1 = ev(), ev(), ev(), ev()
        # interleave the host-side enqueues so that neither stream starts far ahead
        ia = ib = 0
        with torch.cuda.stream(sA):
            a0.record()
        with torch.cuda.stream(sB):
            b0.record()
        while ia < na or ib < nb:
            if fa is not None and ia < na:
                with torch.cuda.stream(sA):
                    fa()
                ia += 1
            elif fa is None:
                ia = na
            if fb is not None and ib * na < ia * nb or ia >= na:
                if fb is not None and ib < nb:
                    with torch.cuda.stream(sB):
                        fb()
                    ib += 1
                elif fb is None:
                    ib = nb
        with torch.cuda.stream(sA):
            a1.record()
        with torch.cuda.stream(sB):
            b1.record()
        torch.cuda.synchronize()
        return a0.elapsed_time(a1), b0.elapsed_time(b1)

    alone = {}
    for k, f in kern.items():
        run(f, 10, None, 0)
        alone[k] = run(f, 60, None, 0)[0] / 60
    print(json.dumps({"schedule": a.schedule, "alone_ms": {k: round(v, 4) for k, v in alone.items()}}), flush=True)
    for ka in ("cv_fwd", "cv_bwd"):
        for kb in ("head_fwd", "head_bwd"):
            na = 60
            nb = max(1, round(na * alone[ka] / alone[kb]))
            run(kern[ka], 10, kern[kb], 10)
            ta, tb = run(kern[ka], na, kern[kb], nb)
            print(json.dumps({"pair": f"{ka} || {kb}", "n": [na, nb], "alone_ms": [round(na * alone[ka], 3), round(nb * alone[kb], 3)],
                              "together_ms": [round(ta, 3), round(tb, 3)], "slowdown": [round(ta / (na * alone[ka]), 3), round(tb / (nb * alone[kb]), 3)],
                              "sharing_efficiency": round((na * alone[ka] + nb * alone[kb]) / max(ta, tb), 3)}), flush=True)


if __name__ == "__main__":
    main()
