"""Schedules of the hot path around the four kernels: two-stream overlap on the device, and the HOST-buffer entry
points (pinned host tensors in, pinned host results out) that bench.py times as ``e2e``.

Host side: ONE coalesced pinned staging buffer per direction and per slot, so a step issues ONE ``cudaMemcpyAsync``
host->device and ONE device->host (the box's PCIe fabric is the end-to-end limit, profiles/r2_h2d_ceiling.md; fewer,
larger copies are what it wants).  The caller fills the staging views in place (``slot.x`` / ``.y`` / ``.cost_lr`` are
views of the pinned buffer: collate straight into them), then ``submit(slot)``.  Copies run on side streams and are
double-buffered against the compute stream: steady-state throughput is max(PCIe time, kernel time) per step.
"""
from __future__ import annotations

import torch

from . import functional as F_

CV_FWD_SHARED = 3    # include/rag_b200.h RAG_CV_FWD_SHARED
CV_BWD_SHARED = 2    # include/rag_b200.h RAG_CV_BWD_SHARED
CV_FWD_SLIM = 5      # RAG_CV_FWD_SLIM
CV_BWD_SLIM = 3      # RAG_CV_BWD_SLIM
HEAD_FWD_SHARED = 4  # RAG_HEAD_FWD_SHARED
HEAD_BWD_SHARED = 3  # RAG_HEAD_BWD_SHARED

# kernel variants of a schedule: (cost-volume forward, head forward, cost-volume backward, head backward); None = default.
# "launch-order": the volume kernels are persistent grids and must be launched before the head kernels they share the SMs
# with (the head grids are not resident at once).  "coresident": all four are persistent grids whose CTAs fit an SM
# TOGETHER -- volume kernel <= a quarter of the register file, three 128-register head CTAs in the rest -- so they share
# the SMs whichever is launched first (free-running streams cannot control that order).
SCHEDULES = {
    "launch-order": (CV_FWD_SHARED, None, CV_BWD_SHARED, None),
    "coresident": (CV_FWD_SLIM, HEAD_FWD_SHARED, CV_BWD_SLIM, HEAD_BWD_SHARED),
    "slim-volume": (CV_FWD_SLIM, None, CV_BWD_SLIM, None),
}


class OverlappedPath:
    """Device-resident schedule of the hot path for a STREAM of batches: the cost volume runs on one
    CUDA stream and the disparity head on another, so the HBM-bound volume kernel of one batch shares
    the SMs with the FP32-bound head kernel of another (in the network the Matching Net sits between the
    two operators of a batch, so what overlaps in deployment is head(batch i) with volume(batch i+1)).

    ``step`` only enqueues; results are ordered on the returned events (or ``join()``).  ``schedule`` picks the kernel
    variants (SCHEDULES): "coresident" (default) runs every kernel as a persistent grid sized so that a volume CTA and three
    head CTAs fit an SM together, so the two streams overlap whichever kernel is launched first; "launch-order" is the
    earlier form (persistent volume kernel launched first, the head's plain grid dispatched beside it).
    """

    schedule = "coresident"

    def __init__(self, maxdisp: int = 192, device: torch.device | str = "cuda", schedule: str | tuple | None = None):
        self.device = torch.device(device)
        self.maxdisp = maxdisp
        self.s_cv = torch.cuda.Stream(self.device)
        self.s_head = torch.cuda.Stream(self.device)
        if schedule is not None:
            self.schedule = schedule
        self.variants = SCHEDULES[self.schedule] if isinstance(self.schedule, str) else tuple(self.schedule)

    def _fork(self):
        cur = torch.cuda.current_stream(self.device)
        self.s_cv.wait_stream(cur)
        self.s_head.wait_stream(cur)
        return cur

    def step(self, x: torch.Tensor, y: torch.Tensor, cost_lr: torch.Tensor, want_stats: bool = False):
        """x, y [B,C,Hf,Wf] and cost_lr [B,1,Dl,Hl,Wl] on the device, ready on the current stream.
        Returns (cost, disp, stats_or_None); cost is valid on ``self.s_cv``, disp/stats on ``self.s_head``."""
        cur = self._fork()
        with torch.no_grad():
            vcf, vhf = self._variants(x, cost_lr)[:2]
            with torch.cuda.stream(self.s_cv):
                cost = F_.cost_volume_forward(x, y, int(self.maxdisp / 3), variant=vcf)
            with torch.cuda.stream(self.s_head):
                disp, stats = F_.disp_head_forward(cost_lr, self.maxdisp, want_stats=want_stats, variant=vhf)
        for t in (x, y):
            t.record_stream(self.s_cv)
        cost_lr.record_stream(self.s_head)
        # the outputs live in the side streams' allocator pools but are consumed on the caller's stream (after
        # join()): tell the caching allocator, so a freed block is not recycled under a running consumer
        for t in (cost, disp, stats):
            if t is not None:
                t.record_stream(cur)
        return cost, disp, stats

    def _variants(self, x, cost_lr):
        """The schedule's kernel variants where their preconditions hold (vector widths, x3 ratio), else the defaults."""
        vcf, vhf, vcb, vhb = self.variants

        def misaligned(t):                          # the vector kernels want 16-byte aligned bases (views may not be)
            ptr = getattr(t, "data_ptr", None)
            return ptr is not None and ptr() % 16 != 0

        if x.shape[-1] % 4 != 0 or misaligned(x):
            vcf = vcb = None
        if cost_lr.shape[-1] % 4 != 0 or self.maxdisp != 3 * cost_lr.shape[2] or misaligned(cost_lr):
            vhf = None
        if self.maxdisp != 3 * cost_lr.shape[2]:
            vhb = None
        return vcf, vhf, vcb, vhb

    def join(self):
        """Make the current stream wait for everything enqueued so far."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.s_cv)
        cur.wait_stream(self.s_head)

    def capture(self, *args, **kw):
        """Capture one ``step(*args)`` + ``join()`` (the two-stream fork/join included) into a CUDA graph over the given
        STATIC input tensors.  Returns ``(graph, outputs)``: refill the inputs in place, ``graph.replay()``, read the
        outputs.  A replay costs one launch on the host instead of ~10 calls per step, which matters for the small training
        kernels (80 us each: the eager two-stream step is host-bound).  Every launch of the library is capturable: work
        counters live in per-launch workspaces that the captured memset nodes re-zero on every replay."""
        cur = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self.step(*args, **kw)             # warm-up outside the capture (cudaFuncSetAttribute, allocator pools)
            self.join()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        self._cv_step_done = None                   # no dependency of the captured step on an event from outside the capture
        with torch.cuda.graph(graph, stream=side):
            outs = self.step(*args, **kw)
            self.join()
        self._cv_step_done = None
        return graph, outs


class OverlappedTrainPath(OverlappedPath):
    """The training step (forward + backward of both operators) on the same two streams: the HBM-bound volume kernels
    (cost-volume forward, then its backward) on one stream as persistent grids, the FP32-bound head kernels (forward,
    then backward) on the other.  Inside a network the partners differ (the Matching Net sits between the operators, and
    its backward between their backwards); what the schedule shows is that the two kinds of kernel do not compete for
    the same resource, so a training step costs ~max(volume stream, head stream) instead of their sum."""

    # The two chains take about the same time, so two FREE-running streams settle into whatever phase the start-up left
    # them in, and the step time depends on it (0.196 - 0.224 ms at B=4 288x576 from box to box and run to run:
    # head_fwd beside cv_fwd and head_bwd beside cv_bwd share an SM better than the crossed pairs, tools/corun_rates.py).
    # phase_lock: the head stream starts step k when the volume stream has finished step k-1 -- one event per step, no
    # join -- which pins the aligned pairing: 0.207 - 0.214 ms on every box (tools/couple_probe.py).
    phase_lock = True

    def step(self, x, y, cost_lr, gcost, gdisp):
        cur = self._fork()
        c = x.shape[1]
        vcf, vhf, vcb, vhb = self._variants(x, cost_lr)
        prev = getattr(self, "_cv_step_done", None)
        with torch.no_grad():
            with torch.cuda.stream(self.s_cv):
                cost = F_.cost_volume_forward(x, y, int(self.maxdisp / 3), variant=vcf)
                gx, gy = F_.cost_volume_backward(gcost, c, variant=vcb)
                if self.phase_lock:
                    self._cv_step_done = torch.cuda.Event()
                    self._cv_step_done.record(self.s_cv)
            with torch.cuda.stream(self.s_head):
                if self.phase_lock and prev is not None:
                    self.s_head.wait_event(prev)
                disp, stats = F_.disp_head_forward(cost_lr, self.maxdisp, want_stats=True, variant=vhf)
                gcl = F_.disp_head_backward(cost_lr, gdisp, disp, stats, self.maxdisp, variant=vhb)
        for t in (x, y, gcost):
            t.record_stream(self.s_cv)
        for t in (cost_lr, gdisp):
            t.record_stream(self.s_head)
        for t in (cost, gx, gy, disp, stats, gcl):
            t.record_stream(cur)
        return cost, disp, gcl, gx, gy


class _Slot:
    """One in-flight step: pinned staging (views into one buffer per direction), device mirrors, events."""

    def __init__(self, shapes_in: dict, shapes_out: dict, device):
        def carve(shapes, pinned):
            n = sum(int(torch.Size(s).numel()) for s in shapes.values())
            buf = torch.empty(n, dtype=torch.float32).pin_memory() if pinned else torch.empty(n, dtype=torch.float32, device=device)
            views, o = {}, 0
            for k, s in shapes.items():
                m = int(torch.Size(s).numel())
                views[k] = buf[o:o + m].view(s)
                o += m
            return buf, views

        self.h_in, self.host = carve(shapes_in, True)        # the caller fills self.host[...] in place
        self.d_in, self.dev = carve(shapes_in, False)
        self.d_out, self.dev_out = carve(shapes_out, False)
        self.h_out, self.result = carve(shapes_out, True)    # valid after done.synchronize()
        self.ready, self.free, self.done = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self.keep = None

    def __getattr__(self, name):           # slot.x, slot.y, slot.cost_lr ... -> pinned input views
        host = self.__dict__.get("host", {})
        if name in host:
            return host[name]
        raise AttributeError(name)

    @property
    def h2d_bytes(self):
        return self.h_in.numel() * 4

    @property
    def d2h_bytes(self):
        return self.h_out.numel() * 4


class HostPipeline:
    """Inference from host buffers: features + matching cost in, disparity out.

        pipe = HostPipeline(192, dev)
        slot = pipe.acquire((B, C, Hf, Wf), (B, 1, Dl, Hl, Wl))    # waits until the slot's previous use has finished
        slot.x[...] = ...; slot.y[...] = ...; slot.cost_lr[...] = ...   # fill the pinned staging views in place
        pipe.submit(slot)                                           # 1 H2D copy, 2 kernels, 1 D2H copy -- all async
        slot.done.synchronize(); slot.result["disp"]                # pinned host [B,3Hl,3Wl]
    """

    def __init__(self, maxdisp: int = 192, device: torch.device | str = "cuda", depth: int = 2):
        self.device = torch.device(device)
        self.maxdisp = maxdisp
        self.depth = depth
        self._copy = torch.cuda.Stream(self.device)
        self._d2h = torch.cuda.Stream(self.device)
        self._slots, self._key, self._n = None, None, 0

    # ---- shapes -----------------------------------------------------------------------------------
    def _shapes(self, feat_shape, cost_shape):
        b, _, dl, hl, wl = cost_shape
        return ({"x": tuple(feat_shape), "y": tuple(feat_shape), "cost_lr": tuple(cost_shape)}, {"disp": (b, 3 * hl, 3 * wl)})

    def acquire(self, feat_shape, cost_shape) -> _Slot:
        key = (tuple(feat_shape), tuple(cost_shape))
        if self._slots is None or self._key != key:
            self.drain()
            sin, sout = self._shapes(feat_shape, cost_shape)
            self._slots = [_Slot(sin, sout, self.device) for _ in range(self.depth)]
            self._key, self._n = key, 0
        slot = self._slots[self._n % self.depth]
        self._n += 1
        slot.done.synchronize()              # host may overwrite the staging buffer only after its last copy was issued AND consumed
        return slot

    # ---- the step ---------------------------------------------------------------------------------
    def _compute(self, slot: _Slot, keep_volume: bool):
        with torch.no_grad():
            cost = F_.cost_volume_forward(slot.dev["x"], slot.dev["y"], int(self.maxdisp / 3))   # stays on the device (Matching Net input)
            F_.disp_head_forward(slot.dev["cost_lr"], self.maxdisp, want_stats=False, out=slot.dev_out["disp"])   # straight into the result buffer
        slot.keep = cost if keep_volume else None

    def submit(self, slot: _Slot, keep_volume: bool = False) -> _Slot:
        compute = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._copy):
            self._copy.wait_event(slot.free)      # the previous user of this slot's device buffers has finished
            slot.d_in.copy_(slot.h_in, non_blocking=True)          # ONE cudaMemcpyAsync host->device
            slot.ready.record(self._copy)
        compute.wait_event(slot.ready)
        compute.wait_event(slot.done)             # the previous result of this slot has left d_out
        self._compute(slot, keep_volume)
        slot.free.record(compute)
        self._d2h.wait_event(slot.free)
        with torch.cuda.stream(self._d2h):
            slot.h_out.copy_(slot.d_out, non_blocking=True)        # ONE cudaMemcpyAsync device->host
            slot.done.record(self._d2h)
        return slot

    def run(self, x_h: torch.Tensor, y_h: torch.Tensor, cost_lr_h: torch.Tensor) -> torch.Tensor:
        """Synchronous convenience call for arbitrary host tensors (packs them into the staging buffer on the CPU):
        returns the pinned host disparity map."""
        slot = self.acquire(x_h.shape, cost_lr_h.shape)
        slot.x.copy_(x_h), slot.y.copy_(y_h), slot.cost_lr.copy_(cost_lr_h)
        self.submit(slot)
        slot.done.synchronize()
        return slot.result["disp"]

    def drain(self):
        for s in self._slots or []:
            s.done.synchronize()


class HostTrainPipeline(HostPipeline):
    """Training step from host buffers: forward + backward of both operators.  Host inputs per step: left/right features,
    matching cost, and the upstream disparity gradient (what the masked loss produces from the ground truth); host
    results: disparity, gradient w.r.t. the matching cost, gradients w.r.t. the two feature maps.  The upstream gradient
    of the VOLUME ([B,2C,Df,Hf,Wf], 453 MB at B=4 288x576) is produced on the device by the Matching Net's backward in any
    deployment and never crosses PCIe: it is passed as a device tensor (``gcost_dev``)."""

    def _shapes(self, feat_shape, cost_shape):
        b, _, dl, hl, wl = cost_shape
        img = (b, 3 * hl, 3 * wl)
        return ({"x": tuple(feat_shape), "y": tuple(feat_shape), "cost_lr": tuple(cost_shape), "gdisp": img},
                {"disp": img, "gcost_lr": tuple(cost_shape), "gx": tuple(feat_shape), "gy": tuple(feat_shape)})

    def submit(self, slot: _Slot, gcost_dev: torch.Tensor = None, keep_volume: bool = False) -> _Slot:   # noqa: D102
        self._gcost = gcost_dev
        return super().submit(slot, keep_volume)

    def _compute(self, slot: _Slot, keep_volume: bool):
        c = slot.dev["x"].shape[1]
        with torch.no_grad():
            cost = F_.cost_volume_forward(slot.dev["x"], slot.dev["y"], int(self.maxdisp / 3))
            disp, stats = F_.disp_head_forward(slot.dev["cost_lr"], self.maxdisp, want_stats=True)
            gcl = F_.disp_head_backward(slot.dev["cost_lr"], slot.dev["gdisp"], disp, stats, self.maxdisp)
            gx, gy = F_.cost_volume_backward(self._gcost, c)
            slot.dev_out["disp"].copy_(disp)
            slot.dev_out["gcost_lr"].copy_(gcl)
            slot.dev_out["gx"].copy_(gx)
            slot.dev_out["gy"].copy_(gy)
        slot.keep = cost if keep_volume else None
