import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200.fused_stem import cv_stem_forward
b, c, hf, wf, md = 8, 12, 160, 320, 192
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(b, c, hf, wf, device="cuda", generator=g); y = torch.randn(b, c, hf, wf, device="cuda", generator=g)
w = torch.randn(12, 24, 3, 3, 3, device="cuda", generator=g) * 0.05
sc = torch.rand(12, device="cuda", generator=g) + 0.5; sh = torch.randn(12, device="cuda", generator=g)
for _ in range(3):
    out = cv_stem_forward(x, y, w, sc, sh, True, md)
torch.cuda.synchronize(); print("ok", float(out.mean()))
