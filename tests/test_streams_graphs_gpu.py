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


@pytest.mark.parametrize("schedule", ["launch-order", "coresident"])
def test_overlapped_path_matches_the_serial_modules(schedule):
    """Two-stream schedule (cost volume on one stream, head on another, many steps in flight): every step's
    outputs are bit-identical to the serial modules' (every launch of the persistent cost-volume kernel gets its own
    freshly zeroed work counter; a stale or shared counter would leave rows unwritten)."""
    from rag_b200 import functional as F_
    from rag_b200.pipeline import OverlappedPath

    g = gen(8)
    md = 96
    op = OverlappedPath(md, schedule=schedule)
    batches = [(randn((2, 12, 9, 64), g).cuda(), randn((2, 12, 9, 64), g).cuda(), randn((2, 1, 32, 9, 64), g).cuda()) for _ in range(4)]
    outs = []
    for rep in range(40):                       # far more launches than counter slots in flight
        x, y, cl = batches[rep % 4]
        outs.append(op.step(x, y, cl, want_stats=True))
    op.join()
    torch.cuda.synchronize()
    for rep, (cost, disp, stats) in enumerate(outs):
        x, y, cl = batches[rep % 4]
        assert torch.equal(cost, F_.cost_volume_forward(x, y, md // 3, variant=0)), f"step {rep}: volume"
        d2, s2 = F_.disp_head_forward(cl, md, True)
        assert torch.equal(disp, d2) and torch.equal(stats, s2), f"step {rep}: head"
    assert torch.equal(outs[0][0].cpu(), O.cost_volume_ref(batches[0][0].cpu(), batches[0][1].cpu(), md))


def test_persistent_cost_volume_under_concurrent_launches():
    """Launches of the persistent kernel on several streams at once each own their work counter."""
    from rag_b200 import functional as F_

    g = gen(9)
    xs = [randn((1, 6, 11, 36), g).cuda() for _ in range(6)]
    ys = [randn((1, 6, 11, 36), g).cuda() for _ in range(6)]
    streams = [torch.cuda.Stream() for _ in range(6)]
    torch.cuda.synchronize()
    outs = [[] for _ in range(6)]
    for rep in range(10):
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                outs[i].append(F_.cost_volume_forward(xs[i], ys[i], 16))
    torch.cuda.synchronize()
    for i in range(6):
        ref = O.cost_volume_ref(xs[i].cpu(), ys[i].cpu(), 48)
        for o in outs[i]:
            assert torch.equal(o.cpu(), ref)


def test_captured_graphs_replayed_beside_eager_launches():
    """the persistent cost-volume kernel takes its work counter from a caller-owned workspace that the
    launcher zeroes on the launch stream, so nothing is shared between launches.  Two graphs (each with its own
    counter) replayed on two streams while eager launches of the same kernel run on a third must all produce the
    oracle's bits, 200 times over (with the round-1 global slot pool a captured slot index collided with eager
    launches and left rows unwritten)."""
    from rag_b200 import functional as F_

    g = gen(21)
    md = 96
    shapes = [(2, 12, 9, 64), (1, 12, 14, 96), (3, 12, 5, 128)]
    data = [(randn(s, g).cuda(), randn(s, g).cuda()) for s in shapes]
    refs = [O.cost_volume_ref(x.cpu(), y.cpu(), md) for x, y in data]
    streams = [torch.cuda.Stream() for _ in range(3)]
    graphs, outs = [], []
    for i in range(2):
        x, y = data[i]
        with torch.cuda.stream(streams[i]):
            F_.cost_volume_forward(x, y, md // 3)               # warm-up (cudaFuncSetAttribute)
        streams[i].synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=streams[i]):
            out = F_.cost_volume_forward(x, y, md // 3)          # persistent kernel + its memset node
        graphs.append(gr)
        outs.append(out)
    torch.cuda.synchronize()
    x2, y2 = data[2]
    for it in range(200):
        for i in range(2):
            outs[i].fill_(float("nan"))
        torch.cuda.synchronize()
        for i in range(2):
            with torch.cuda.stream(streams[i]):
                graphs[i].replay()
        with torch.cuda.stream(streams[2]):
            eager = [F_.cost_volume_forward(x2, y2, md // 3) for _ in range(3)]
        torch.cuda.synchronize()
        for i in range(2):
            assert torch.equal(outs[i].cpu(), refs[i]), f"iteration {it}: graph {i}"
        for e in eager:
            assert torch.equal(e.cpu(), refs[2]), f"iteration {it}: eager"


def test_concurrent_stem_and_last_conv_launches_with_different_weights():
    """more launches in flight than the old 4 / 8 weight slots, different weights, two streams.  The stem's
    regrouped weights now live in a per-launch caller workspace; the last conv's constant-memory slots are ordered with
    events.  Every result must equal the one computed alone."""
    from rag_b200.fused_stem import cv_stem_forward
    from rag_b200.last_conv import conv3d_c1_forward

    g = gen(22)
    md = 48
    x, y = randn((2, 12, 12, 64), g).cuda(), randn((2, 12, 12, 64), g).cuda()
    feat = randn((2, 12, 16, 24, 64), g).cuda()
    ws = [randn((12, 24, 3, 3, 3), g).cuda() for _ in range(12)]
    wl = [randn((1, 12, 3, 3, 3), g).cuda() for _ in range(20)]
    alone_s = [cv_stem_forward(x, y, w, maxdisp=md) for w in ws]
    alone_l = [conv3d_c1_forward(feat, w) for w in wl]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for rep in range(5):
        got_s, got_l = [], []
        for i, w in enumerate(ws):
            with torch.cuda.stream(streams[i % 2]):
                got_s.append(cv_stem_forward(x, y, w, maxdisp=md))
        for i, w in enumerate(wl):
            with torch.cuda.stream(streams[(i + rep) % 2]):
                got_l.append(conv3d_c1_forward(feat, w))
        torch.cuda.synchronize()
        for a, b in zip(got_s, alone_s):
            assert torch.equal(a, b), f"rep {rep}: stem"
        for a, b in zip(got_l, alone_l):
            assert torch.equal(a, b), f"rep {rep}: last conv"


def test_last_conv_refuses_capture_and_the_module_path_falls_back():
    """rag_conv3d_c1_fwd cannot be captured (event-ordered constant slots): the raw call reports RAG_E_CAPTURE, the
    ConvBR_3d drop-in takes the reference's own nn.Conv3d while capturing."""
    from rag_b200 import last_conv as LC

    g = gen(23)
    feat = randn((1, 12, 8, 16, 32), g).cuda()
    conv = torch.nn.Conv3d(12, 1, 3, padding=1, bias=False).cuda()
    with torch.no_grad():
        want = LC.conv_forward(conv, feat)
        conv(feat)                                               # cuDNN warm-up outside the capture
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            assert not LC.qualifies(conv, feat)
            out = LC.conv_forward(conv, feat)
            with pytest.raises(RuntimeError, match="capturing"):
                LC.conv3d_c1_forward(feat, conv.weight)
        gr.replay()
        torch.cuda.synchronize()
    assert (out - want).abs().max().item() <= 1e-3 * want.abs().max().item()
