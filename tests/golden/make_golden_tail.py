"""Golden vectors for the Matching Net's TAIL in training, from the unmodified
reference running on CPU:  the tail of ``Network.matching`` (rag_model.py:356-366) on a level-12 feature map

    mat = last_3_3d( upsample_6( last_6_3d( upsample_12( last_12_3d(x) ) ) ) )

is driven through the reference's own method with hooks on its ``nn.Upsample`` calls and on ``last_3_3d``: for each of the
two upsamples and for the last convolution the file holds input, output, upstream gradient and the autograd gradients
(w.r.t. the input; for the convolution also w.r.t. its weight).

    python tests/golden/make_golden_tail.py        (build container only: imports /root/reference/src)
"""
from __future__ import annotations

import os

import numpy as np
import torch

from make_golden import HERE, _import_reference


def main():
    rm, _ = _import_reference()
    from automl.operations_3d import ConvBR_3d

    g = torch.Generator().manual_seed(1357)
    torch.manual_seed(5)
    c = 12
    d, h, w = 16, 12, 24                                  # the volume's (Df, Hf, Wf); the level-12 map is a quarter of it
    rec = {}

    class Tap(torch.nn.Upsample):                          # records what the reference's nn.Upsample sees and returns
        def forward(self, x):
            x = x.detach().clone().requires_grad_(True)   # cut the graph: each piece gets its own golden gradient
            out = super().forward(x)
            rec.setdefault("ups", []).append((x, out))
            return out

    stub = torch.nn.Module()
    stub.stem3d0 = torch.nn.ModuleList([torch.nn.Identity()])
    stub.stem3d1 = torch.nn.ModuleList([torch.nn.Identity()])
    stub.cells_3d = []
    stub.last_3_3d = torch.nn.ModuleList([ConvBR_3d(c, 1, 3, 1, 1, bn=False, relu=False)])
    stub.last_6_3d = torch.nn.ModuleList([ConvBR_3d(2 * c, c, 1, 1, 0)])
    stub.last_12_3d = torch.nn.ModuleList([ConvBR_3d(4 * c, 2 * c, 1, 1, 0)])
    # matching() takes x = the volume (only its size is used by the tail) and runs stem/cells first; feed it a level-12 map by
    # making stem3d0/stem3d1 identities on a tensor that carries the volume's size through a wrapper
    lvl12 = torch.randn(1, 4 * c, d // 4, h // 4, w // 4, generator=g)

    class Vol:                                             # what matching() calls .size() on
        def size(self):
            return torch.Size((1, 2 * c, d, h, w))

    stub.stem3d0 = torch.nn.ModuleList([_Const(lvl12)])
    stub.stem3d1 = torch.nn.ModuleList([_Const(lvl12)])
    real_nn = rm.nn

    class Proxy:
        Upsample = Tap

        def __getattr__(self, k):
            return getattr(real_nn, k)

    seen = {}
    hook = stub.last_3_3d[0].register_forward_hook(lambda m, i, o: seen.update(x=i[0], out=o))
    rm.nn = Proxy()
    try:
        arch = {k: [0] for k in ("stem_3d0", "stem_3d1", "last_3_3d", "last_6_3d", "last_12_3d")}
        mat = rm.Network.matching(stub, Vol(), arch, None)
    finally:
        rm.nn = real_nn
        hook.remove()
    assert mat.shape == (1, 1, d, h, w) and len(rec["ups"]) == 2
    out = {}
    for name, (x, o) in zip(("up12", "up6"), rec["ups"]):
        go = torch.randn(o.shape, generator=g)
        (gx,) = torch.autograd.grad(o, x, go, retain_graph=True)
        out.update({f"{name}_in": x.detach().numpy(), f"{name}_out": o.detach().numpy(), f"{name}_gout": go.numpy(), f"{name}_gin": gx.numpy()})
    conv = stub.last_3_3d[0].conv
    xin = seen["x"].detach().clone().requires_grad_(True)
    y = stub.last_3_3d[0](xin)
    gy = torch.randn(y.shape, generator=g)
    gxin, gwt = torch.autograd.grad(y, (xin, conv.weight), gy)
    out.update(conv_in=xin.detach().numpy(), conv_weight=conv.weight.detach().numpy(), conv_out=y.detach().numpy(), conv_gout=gy.numpy(),
               conv_gin=gxin.numpy(), conv_gweight=gwt.numpy())
    np.savez_compressed(os.path.join(HERE, "tail_c12_d16_h12_w24.npz"), **out)
    print({k: v.shape for k, v in out.items()})


class _Const(torch.nn.Module):
    def __init__(self, t):
        super().__init__()
        self.t = t

    def forward(self, _):
        return self.t


if __name__ == "__main__":
    main()
