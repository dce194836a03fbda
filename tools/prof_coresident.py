"""Launch every kernel of the co-resident schedule once at the training workload, then the default volume forward at the
inference workload (for ncu captures).

    python tools/prof_coresident.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200 import functional as F_  # noqa: E402
from rag_b200 import pipeline as P_  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1234)
for (b, hf, wf, train) in ((4, 96, 192, True), (8, 160, 320, False)):
    c, df, md = 12, 64, 192
    x = torch.randn(b, c, hf, wf, device="cuda", generator=g)
    y = torch.randn(b, c, hf, wf, device="cuda", generator=g)
    cost = F_.cost_volume_forward(x, y, df)                                   # default = RAG_CV_FWD_SLIM geometry
    if train:
        cl = torch.randn(b, 1, df, hf, wf, device="cuda", generator=g)
        gd = torch.randn(b, 3 * hf, 3 * wf, device="cuda", generator=g) * (torch.rand(b, 3 * hf, 3 * wf, device="cuda", generator=g) < 0.3)
        gc = torch.randn(b, 2 * c, df, hf, wf, device="cuda", generator=g)
        disp, stats = F_.disp_head_forward(cl, md, True, variant=P_.HEAD_FWD_SHARED)
        gcl = F_.disp_head_backward(cl, gd, disp, stats, md, variant=P_.HEAD_BWD_SHARED)
        gx, gy = F_.cost_volume_backward(gc, c, variant=P_.CV_BWD_SLIM)
        gx0, gy0 = F_.cost_volume_backward(gc, c)
    torch.cuda.synchronize()
    del cost
print("ok")
