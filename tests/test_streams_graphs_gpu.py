"""Stream / CUDA-graph / device-guard behaviour of the C-ABI launches."""
import pytest
import torch

from oracle import rag_oracle as O
from tests._util import gen, randn

pytestmark = pytest.mark.gpu


def test_cuda_graph_capture_and_replay():
    """The whole hot path (4 kernels + combine) is capturable: no host sync, no allocation by the library,
    launches on the capturing stream.  Replays must give the same bits as eager execution."""
    from rag_b200 import functional as F_

    g = gen(5)
    md = 96
    x, y = randn((2, 12, 8, 64), g).cuda(), randn((2, 12, 8, 64), g).cuda()
    cl = randn((2, 1, 32, 8, 64), g).cuda()
    gd = randn((2, 24, 192), g).cuda()
    gc = randn((2, 24, 32, 8, 64), g).cuda()

    def run():
        cost = F_.cost_volume_forward(x, y, md // 3)
        disp, stats = F_.disp_head_forward(cl, md, True)
        gcl = F_.disp_head_backward(cl, gd, disp, stats, md)
        gx, gy = F_.cost_volume_backward(gc, 12)
        return cost, disp, gcl, gx, gy

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()  # warm-up on the side stream (cudaFuncSetAttribute etc.)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs = run()
    for _ in range(3):
        x.add_(0.5)   # change inputs in place: every replay must see the new values
        cl.mul_(1.25)
        graph.replay()
    torch.cuda.synchronize()
    replayed = [t.clone() for t in outs]
    eager = run()
    torch.cuda.synchronize()
    for a, b in zip(replayed, eager):
        assert torch.equal(a, b)


def test_launch_respects_the_current_stream():
    """A kernel queued behind a long-running producer on a side stream must see the producer's data."""
    from rag_b200 import functional as F_

    g = gen(6)
    x0, y0 = randn((1, 12, 16, 128), g).cuda(), randn((1, 12, 16, 128), g).cuda()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        x = x0
        for _ in range(50):
            x = x * 1.0001          # keep the side stream busy; x is produced late
        cost = F_.cost_volume_forward(x, y0, 16)
    side.synchronize()
    assert torch.equal(cost.cpu(), O.cost_volume_ref(x.cpu(), y0.cpu(), 48))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_non_current_device():
    """Tensors on cuda:1 while cuda:0 is current (one process driving replicas): the wrapper switches device."""
    from rag_b200.modules import CostVolume, Disp

    g = gen(7)
    x, y = randn((1, 12, 6, 32), g), randn((1, 12, 6, 32), g)
    cl = randn((1, 1, 16, 6, 32), g)
    torch.cuda.set_device(0)
    cost = CostVolume(48)(x.to("cuda:1"), y.to("cuda:1"))
    disp = Disp(48)(cl.to("cuda:1"))
    assert cost.device.index == 1 and disp.device.index == 1
    assert torch.equal(cost.cpu(), O.cost_volume_ref(x, y, 48))
    assert (disp.cpu() - O.disp_head_ref(cl, 48)).abs().max().item() <= 1e-4
    with pytest.raises(RuntimeError):
        CostVolume(48)(x.to("cuda:0"), y.to("cuda:1"))
