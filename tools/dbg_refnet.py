import sys, os, torch
from copy import deepcopy
sys.path.insert(0, '.')
from oracle import refimport as R
from rag_b200 import network as N
import tests.test_reference_network_gpu as T
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
ref = R.import_reference()
torch.manual_seed(3)
net = ref.rag_model.Network(R.make_genotype(ref, 3), "cuda").cuda().train()
left, right, w = T._inputs(2, 6)
call = lambda: net.forward(left, right, 0, net.arch_init)
T._calibrate(net, call, [net.last_3_3d[0]])
state = deepcopy(net.state_dict())
def run():
    net.load_state_dict(state)
    return T._fwd_bwd(net, call, w)
N.uninstall()
ro, rg = run()
ro2, rg2 = run()
def rep(tag, go, gg):
    worst = max(((gg[n] - rg[n]).norm() / rg[n].norm().clamp_min(1e-30)).item() for n in rg)
    wn = max(rg, key=lambda n: ((gg[n] - rg[n]).norm() / rg[n].norm().clamp_min(1e-30)).item())
    print(f"{tag:28s} out max {(go-ro).abs().max().item():.2e}  worst grad rel-L2 {worst:.2e} ({wn})", flush=True)
rep("reference again", ro2, rg2)
for tag, kw in (("core", {}), ("core+upsample", dict(upsample=True)), ("core+fuse_stem", dict(operations_3d=ref.operations_3d, fuse_stem=True)),
                ("all", dict(operations_3d=ref.operations_3d, fuse_stem=True, upsample=True))):
    N.uninstall()
    N.install(ref.rag_model, ref.mdenas_basicmodel, **kw)
    go, gg = run()
    rep(tag, go, gg)
# fuse_stem but last conv through cudnn: disable conv kernel
import rag_b200.last_conv as LC
orig = LC._layer_ok
LC._layer_ok = lambda conv, x: False
N.uninstall(); N.install(ref.rag_model, ref.mdenas_basicmodel, operations_3d=ref.operations_3d, fuse_stem=True)
go, gg = run(); rep("fuse_stem w/o last-conv kernel", go, gg)
LC._layer_ok = orig
import rag_b200.fused_stem as FS
orig2 = FS._train_fusable
FS._train_fusable = lambda self, vol: False
N.uninstall(); N.install(ref.rag_model, ref.mdenas_basicmodel, operations_3d=ref.operations_3d, fuse_stem=True)
go, gg = run(); rep("last-conv kernel w/o fused stem", go, gg)
# ---- fp64 truth of the unpatched network ----
N.uninstall()
FS._train_fusable = orig2
net64 = deepcopy(net).double()
net64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in state.items()})
l64, r64, w64 = left.double(), right.double(), w.double()
net64.zero_grad(set_to_none=True)
o64 = net64.forward(l64, r64, 0, net64.arch_init)
(o64 * w64).sum().backward()
g64 = {n: p.grad for n, p in net64.named_parameters() if p.grad is not None}
def rep64(tag, gg):
    errs = {n: ((gg[n].double() - g64[n]).norm() / g64[n].norm().clamp_min(1e-30)).item() for n in g64}
    wn = max(errs, key=errs.get)
    import statistics
    print(f"{tag:28s} vs fp64 network: worst rel-L2 {errs[wn]:.2e} ({wn}), median {statistics.median(errs.values()):.2e}", flush=True)
rep64("reference fp32", rg)
N.install(ref.rag_model, ref.mdenas_basicmodel, operations_3d=ref.operations_3d, fuse_stem=True, upsample=True)
go, gg = run(); rep64("patched (all)", gg)
N.uninstall(); N.install(ref.rag_model, ref.mdenas_basicmodel)
go, gg = run(); rep64("patched (core)", gg)
