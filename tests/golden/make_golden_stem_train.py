"""Golden vectors for the TRAINING-mode stem ("and its mirror in backward"):
the reference's own stem3d0 = ConvBR_3d(2C, C, 3, 1, 1) (rag_model.py:234,341; operations_3d.py:31-47) in
train() -- BatchNorm3d on batch statistics -- applied to the volume the reference's Network.forward builds
(rag_model.py:375-383), with autograd gradients w.r.t. the two feature maps, the conv weight and the
BatchNorm affine parameters, and the running statistics after the step.

    python tests/golden/make_golden_stem_train.py        (build container only: imports /root/reference/src)

Separate from make_golden.py so that adding it did not touch the existing fixtures.  No kernel consumes
these yet: they pin the oracle (oracle.rag_oracle.stem_ref) for the next round's fused training path.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from make_golden import HERE, _import_reference, _ref_cost_volume


def main():
    rm, _ = _import_reference()
    from automl.operations_3d import ConvBR_3d

    g = torch.Generator().manual_seed(2468)
    torch.manual_seed(77)
    b, c, hf, wf, md = 2, 12, 5, 12, 24
    x = torch.randn(b, c, hf, wf, generator=g).requires_grad_(True)
    y = torch.randn(b, c, hf, wf, generator=g).requires_grad_(True)
    stem = ConvBR_3d(2 * c, c, 3, 1, 1).train()
    with torch.no_grad():
        stem.bn.running_mean.copy_(0.3 * torch.randn(c, generator=g))
        stem.bn.running_var.copy_(0.5 + torch.rand(c, generator=g))
        stem.bn.weight.copy_(0.5 + torch.rand(c, generator=g))
        stem.bn.bias.copy_(0.2 * torch.randn(c, generator=g))
    mean0, var0 = stem.bn.running_mean.clone(), stem.bn.running_var.clone()
    vol = _ref_cost_volume(rm, x, y, md)
    out = stem(vol)
    gout = torch.randn(out.shape, generator=g)
    out.backward(gout)
    np.savez_compressed(
        os.path.join(HERE, "trainstem_b2_c12_h5_w12_md24.npz"),
        x=x.detach().numpy(), y=y.detach().numpy(), maxdisp=md,
        weight=stem.conv.weight.detach().numpy(), bn_weight=stem.bn.weight.detach().numpy(), bn_bias=stem.bn.bias.detach().numpy(),
        bn_mean0=mean0.numpy(), bn_var0=var0.numpy(), bn_eps=stem.bn.eps, bn_momentum=stem.bn.momentum,
        out=out.detach().numpy(), gout=gout.numpy(), gx=x.grad.numpy(), gy=y.grad.numpy(),
        gweight=stem.conv.weight.grad.numpy(), gbn_weight=stem.bn.weight.grad.numpy(), gbn_bias=stem.bn.bias.grad.numpy(),
        bn_mean1=stem.bn.running_mean.numpy(), bn_var1=stem.bn.running_var.numpy(),
    )
    print("stem_train", tuple(out.shape))


if __name__ == "__main__":
    main()
