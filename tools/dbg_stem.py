import sys, copy, torch
sys.path.insert(0, '.')
from tests.test_fused_stem_gpu import ConvBR_3d, ref_stem, _mx
from tests._util import gen, randn
from rag_b200.fused_stem import VirtualCostVolume, stem_forward
torch.backends.cudnn.allow_tf32 = False
for (b, hf, wf, md, o) in [(1, 96, 192, 192, 12), (4, 96, 192, 192, 12), (4, 24, 192, 192, 12), (1, 96, 64, 192, 12)]:
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
        out.backward(gout.to(dtype))
        return [out.detach(), x.grad, y.grad, lay.conv.weight.grad, lay.bn.weight.grad, lay.bn.bias.grad]
    r64 = run(lambda lay, x, y: ref_stem(x, y, lay, md), torch.float64)
    r32 = run(lambda lay, x, y: ref_stem(x, y, lay, md), torch.float32)
    ours = run(lambda lay, x, y: stem_forward(lay, VirtualCostVolume(x, y, md)), torch.float32)
    print((b, hf, wf, md, o), ' '.join(f"{n}: ours {_mx(a, c):.1e} ref32 {_mx(r, c):.1e} |" for n, a, r, c in zip(["out", "gx", "gy", "gw", "gg", "gb"], ours, r32, r64)), flush=True)
    del r64, r32, ours
    torch.cuda.empty_cache()
