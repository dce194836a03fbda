"""Launch cost-volume forward variants (ncu capture helper).  usage: prof_cv.py [v ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200 import functional as F_
b, c, hf, wf, df = 8, 12, 160, 320, 64
g = torch.Generator(device="cuda").manual_seed(1234)
x = torch.randn(b, c, hf, wf, device="cuda", generator=g)
y = torch.randn(b, c, hf, wf, device="cuda", generator=g)
vs = [int(a) for a in sys.argv[1:]] or [18]
for v in vs:
    for _ in range(2):
        out = F_.cost_volume_forward(x, y, df, variant=v)
torch.cuda.synchronize()
print("ok", float(out[0, 0, 0].mean()))
