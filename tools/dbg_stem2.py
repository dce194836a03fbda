import sys, copy, torch
sys.path.insert(0, '.')
from tests.test_fused_stem_gpu import ConvBR_3d, ref_stem_masked, _mx
from tests._util import gen, randn
from rag_b200.fused_stem import VirtualCostVolume, stem_forward
torch.backends.cudnn.allow_tf32 = False
for (b, hf, wf, md, o) in [(2, 32, 64, 192, 12), (2, 32, 64, 96, 12), (2, 32, 60, 192, 12), (1, 8, 64, 192, 12)]:
    g = gen(6)
    layer = ConvBR_3d(24, o).cuda().train()
    x0, y0 = randn((b, 12, hf, wf), g).cuda(), randn((b, 12, hf, wf), g).cuda()
    gout = randn((b, o, md // 3, hf, wf), g).cuda()
    state = copy.deepcopy(layer.state_dict())
    def run(fn, dtype):
        lay = copy.deepcopy(layer).to(dtype)
        lay.load_state_dict({k: v.to(dtype) if v.is_floating_point() else v for k, v in state.items()})
        x, y = x0.to(dtype).clone().requires_grad_(True), y0.to(dtype).clone().requires_grad_(True)
        out = fn(lay, x, y)
        if isinstance(out, tuple): out = out[0]
        out.backward(gout.to(dtype))
        return [out.detach(), x.grad, y.grad, lay.conv.weight.grad, lay.bn.weight.grad, lay.bn.bias.grad]
    ours = run(lambda lay, x, y: stem_forward(lay, VirtualCostVolume(x, y, md)), torch.float32)
    mask = ours[0] > 0
    r64 = run(lambda lay, x, y: ref_stem_masked(x, y, lay, md, mask), torch.float64)
    print((b, hf, wf, md, o), ' '.join(f"{n}: {_mx(a, c):.1e} sum-rel {abs((a.double().sum()-c.sum()).item())/max(abs(c.sum().item()),1e-30):.1e} |" for n, a, c in zip(["out", "gx", "gy", "gw", "gg", "gb"], ours, r64)), flush=True)
    # where is the gx error
    e = (ours[1].double() - r64[1]).abs()
    print('   gx err by column (max over rest):', [f"{v:.1e}" for v in e.amax(dim=(0,1,2))[:6].tolist()], '...', [f"{v:.1e}" for v in e.amax(dim=(0,1,2))[-4:].tolist()])
    e = (ours[2].double() - r64[2]).abs()
    print('   gy err by column:', [f"{v:.1e}" for v in e.amax(dim=(0,1,2))[:6].tolist()], '...', [f"{v:.1e}" for v in e.amax(dim=(0,1,2))[-4:].tolist()])
