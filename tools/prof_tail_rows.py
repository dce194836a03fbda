"""Launch the trilinear row kernels and the last_3_3d backward kernels once at the training shape (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_b200 import _cabi  # noqa: E402
from rag_b200 import upsample as U  # noqa: E402

L = _cabi.lib()
g = torch.Generator(device="cuda").manual_seed(2)
st = torch.cuda.current_stream().cuda_stream
shape, size = (4, 12, 32, 48, 96), (64, 96, 192)
x = torch.randn(*shape, device="cuda", generator=g)
go = torch.randn(*shape[:2], *size, device="cuda", generator=g)
for _ in range(2):
    U._resize("rag_trilinear_resize_fwd", x, tuple(shape[:2]) + size, tuple(shape[2:]), size, True)
    U._resize("rag_trilinear_resize_bwd", go, tuple(shape), tuple(shape[2:]), size, True)
b, c, d, h, w = 4, 12, 64, 96, 192
xi = torch.randn(b, c, d, h, w, device="cuda", generator=g)
wt = torch.randn(1, c, 3, 3, 3, device="cuda", generator=g) * 0.1
g1 = torch.randn(b, 1, d, h, w, device="cuda", generator=g)
gin, gw = torch.empty_like(xi), torch.empty_like(wt)
ws = torch.empty(int(L.rag_conv3d_c1_bwd_workspace_bytes(c)) // 4, device="cuda")
for _ in range(2):
    _cabi.check(L.rag_conv3d_c1_bwd(g1.data_ptr(), xi.data_ptr(), wt.data_ptr(), gin.data_ptr(), gw.data_ptr(), ws.data_ptr(), b, c, d, h, w, st), "bwd")
torch.cuda.synchronize()
print("ok")
