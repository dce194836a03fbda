"""2+ rank NCCL smoke of the training-side plumbing (launch with torchrun): pairs sharded across ranks,
hot path through the drop-in forward, fused masked loss, DDP gradient all-reduce, metric all-reduce.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_smoke.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200 import dist as D  # noqa: E402
from rag_b200 import metrics as M  # noqa: E402
from rag_b200 import network as N  # noqa: E402
from rag_b200.modules import Disp  # noqa: E402
from tests._mirror_net import MirrorNet, arch  # noqa: E402


class Wrapped(torch.nn.Module):
    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, left, right, t, task_arch):
        return N.network_forward(self.net, left, right, t, task_arch)


def main():
    rank, world, local = D.init()
    dev = torch.device("cuda", local)
    torch.manual_seed(0)                                   # same initial weights on every rank
    md = 48
    net = MirrorNet(Disp(md), md, n_paths=2).to(dev)
    model = D.wrap_ddp(Wrapped(net), local)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9)
    g = torch.Generator().manual_seed(123)
    n_pairs = 8
    left, right = torch.randn(n_pairs, 3, 36, 72, generator=g), torch.randn(n_pairs, 3, 36, 72, generator=g)
    gt = torch.rand(n_pairs, 36, 72, generator=g) * 40
    mine = list(D.shard_pairs(n_pairs, rank, world))
    l, r, t = left[mine].to(dev), right[mine].to(dev), gt[mine].to(dev)
    acc = M.MetricAccumulator(dev)
    losses = []
    for step in range(5):
        opt.zero_grad()
        disp = model(l, r, 0, arch(0))
        loss, sums = M.masked_smooth_l1(disp, t, md)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)   # approaches/rag.py:215
        opt.step()
        acc.update(sums)
        losses.append(loss.item())
    # gradients / weights must be identical on every rank after DDP's all-reduce
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(flat, ref), "weights diverged across ranks"
    out = acc.mean()
    if rank == 0:
        print("ddp_smoke ok: world", world, "losses", [round(x, 4) for x in losses], "metrics", {k: round(v, 4) for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
