"""Launch the last_3_3d conv kernel (ncu capture helper)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200.last_conv import conv3d_c1_forward
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(8, 12, 64, 160, 320, device="cuda", generator=g)
w = torch.randn(1, 12, 3, 3, 3, device="cuda", generator=g) * 0.1
for _ in range(3):
    o = conv3d_c1_forward(x, w)
torch.cuda.synchronize()
print("ok", float(o.mean()))
