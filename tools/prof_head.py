"""Launch each x3 head-forward variant twice (ncu capture helper)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200 import functional as F_
b, hf, wf, df, md = 8, 160, 320, 64, 192
g = torch.Generator(device="cuda").manual_seed(1234)
cl = torch.randn(b, 1, df, hf, wf, device="cuda", generator=g)
for v in (6, 6):
    d, s = F_.disp_head_forward(cl, md, True, variant=v)
torch.cuda.synchronize()
print("ok", float(d.mean()))
