import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_b200 import _cabi
L = _cabi.lib()
g = torch.Generator(device="cuda").manual_seed(5)
b, c, d, h, w = 8, 12, 64, 160, 320
x = torch.randn(b, c, d, h, w, device="cuda", generator=g)
wt = torch.randn(1, c, 3, 3, 3, device="cuda", generator=g) * 0.1
out = torch.empty(b, 1, d, h, w, device="cuda")
for _ in range(2):
    _cabi.check(L.rag_conv3d_c1_fwd(x.data_ptr(), wt.data_ptr(), out.data_ptr(), b, c, d, h, w, torch.cuda.current_stream().cuda_stream), "fwd")
torch.cuda.synchronize(); print("ok")
