import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_b200 import fused_stem as FS
g = torch.Generator(device="cuda").manual_seed(3)
b, hf, wf, md, o = 8, 160, 320, 192, 12
x = torch.randn(b, 12, hf, wf, device="cuda", generator=g); y = torch.randn(b, 12, hf, wf, device="cuda", generator=g)
w = torch.randn(o, 24, 3, 3, 3, device="cuda", generator=g) * 0.1
sc = torch.rand(o, device="cuda", generator=g) + 0.5; sh = torch.randn(o, device="cuda", generator=g)
for _ in range(2):
    out = FS.cv_stem_forward(x, y, w, sc, sh, True, md)
torch.cuda.synchronize(); print("ok")
