"""Launch head-forward variants (ncu capture helper).  usage: prof_head.py [v ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200 import functional as F_
b, hf, wf, df, md = 8, 160, 320, 64, 192
g = torch.Generator(device="cuda").manual_seed(1234)
cl = torch.randn(b, 1, df, hf, wf, device="cuda", generator=g)
vs = [int(a) for a in sys.argv[1:]] or [10]
for v in vs:
    for _ in range(2):
        d, s = F_.disp_head_forward(cl, md, True, variant=v)
torch.cuda.synchronize()
print("ok", float(d.mean()))
