"""Launch each default hot-path kernel twice at one workload (for ncu captures).

    python tools/prof_once.py [infer|train]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200 import functional as F_  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "infer"
b, c, hf, wf, df, md = (8, 12, 160, 320, 64, 192) if wl == "infer" else (4, 12, 96, 192, 64, 192)
g = torch.Generator(device="cuda").manual_seed(1234)
x = torch.randn(b, c, hf, wf, device="cuda", generator=g)
y = torch.randn(b, c, hf, wf, device="cuda", generator=g)
cl = torch.randn(b, 1, df, hf, wf, device="cuda", generator=g)
gd = torch.randn(b, 3 * hf, 3 * wf, device="cuda", generator=g) * (torch.rand(b, 3 * hf, 3 * wf, device="cuda", generator=g) < 0.3)
gc = torch.randn(b, 2 * c, df, hf, wf, device="cuda", generator=g)
for _ in range(2):
    cost = F_.cost_volume_forward(x, y, df)
    disp, stats = F_.disp_head_forward(cl, md, True)
    gcl = F_.disp_head_backward(cl, gd, disp, stats, md)
    gx, gy = F_.cost_volume_backward(gc, c)
torch.cuda.synchronize()
print("ok", float(disp.mean()), float(gcl.abs().mean()))
