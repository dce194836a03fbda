"""Small shapes through every kernel/variant (for compute-sanitizer memcheck/racecheck runs)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200 import functional as F_, metrics as M, staging as S

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
for (b, c, hf, wf, md) in [(1, 12, 9, 36, 96), (2, 3, 5, 37, 30), (1, 4, 7, 70, 192), (1, 2, 3, 10, 48)]:
    x = torch.randn(b, c, hf, wf, device=dev, generator=g); y = torch.randn(b, c, hf, wf, device=dev, generator=g)
    for v in (0, 1, 2, 3):
        cost = F_.cost_volume_forward(x, y, int(md / 3), variant=v)
    gc = torch.randn_like(cost)
    for v in (0, 1):
        F_.cost_volume_backward(gc, c, variant=v)
for (b, dl, hl, wl, md) in [(1, 16, 5, 36, 48), (2, 8, 3, 7, 24), (1, 20, 9, 4, 60), (1, 12, 3, 4, 48)]:
    cl = torch.randn(b, 1, dl, hl, wl, device=dev, generator=g)
    gd = torch.randn(b, 3 * hl, 3 * wl, device=dev, generator=g)
    x3 = md == 3 * dl
    for vf in [0] + ([1, 2, 3] if x3 else []) + ([4, 5, 6, 7] if x3 and wl % 4 == 0 else []):
        disp, st = F_.disp_head_forward(cl, md, True, variant=vf)
    for vb in [0] + ([1, 2] if x3 else []):
        F_.disp_head_backward(cl, gd, disp, st, md, variant=vb)
    F_.upsample_trilinear(cl, md, True)
p = torch.softmax(torch.randn(1, 24, 6, 10, device=dev, generator=g), 1).contiguous().requires_grad_(True)
F_.DisparityRegressionFn.apply(p, 24).sum().backward()
gt = torch.rand(2, 20, 30, device=dev, generator=g) * 250 - 20
est = (gt + torch.randn(2, 20, 30, device=dev, generator=g)).requires_grad_(True)
loss, sums = M.masked_smooth_l1(est, gt, 192); loss.backward()
S.normalize_pad(torch.randint(0, 256, (2, 10, 14, 3), device=dev, dtype=torch.uint8), 12, 18)
torch.cuda.synchronize()
print("sanitize_smoke ok")
