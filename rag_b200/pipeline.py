"""Host-buffer entry point of the hot path: pinned host tensors in, pinned host disparity out.

This is the call a host-side user of the path makes when the data does not already live on the
GPU (and what bench.py times as ``e2e``): per step it copies the step's inputs host->device,
runs the cost-volume and disparity-head kernels through the public modules, and copies the
disparity map device->host.  Copies run on a side stream and are double-buffered against the
compute stream, so steady-state throughput is max(PCIe time, kernel time) per step.
"""
from __future__ import annotations

import torch

from . import functional as F_
from .modules import CostVolume, Disp

CV_FWD_SHARED = 3    # include/rag_b200.h RAG_CV_FWD_SHARED


class OverlappedPath:
    """Device-resident schedule of the hot path for a STREAM of batches: the cost volume runs on one
    CUDA stream and the disparity head on another, so the HBM-bound volume kernel of one batch shares
    the SMs with the FP32-bound head kernel of another (in the network the Matching Net sits between the
    two operators of a batch, so what overlaps in deployment is head(batch i) with volume(batch i+1)).

    ``step`` only enqueues; results are ordered on the returned events (or ``join()``).  The volume kernel
    is launched first with the SM-sharing variant (a persistent grid that is resident at once), which is
    what lets the head's CTAs be dispatched next to it instead of behind it.
    """

    def __init__(self, maxdisp: int = 192, device: torch.device | str = "cuda"):
        self.device = torch.device(device)
        self.maxdisp = maxdisp
        self.s_cv = torch.cuda.Stream(self.device)
        self.s_head = torch.cuda.Stream(self.device)

    def step(self, x: torch.Tensor, y: torch.Tensor, cost_lr: torch.Tensor, want_stats: bool = False):
        """x, y [B,C,Hf,Wf] and cost_lr [B,1,Dl,Hl,Wl] on the device, ready on the current stream.
        Returns (cost, disp, stats_or_None); cost is valid on ``self.s_cv``, disp/stats on ``self.s_head``."""
        cur = torch.cuda.current_stream(self.device)
        self.s_cv.wait_stream(cur)
        self.s_head.wait_stream(cur)
        with torch.no_grad():
            with torch.cuda.stream(self.s_cv):
                cost = F_.cost_volume_forward(x, y, int(self.maxdisp / 3), variant=CV_FWD_SHARED if x.shape[-1] % 4 == 0 else None)
            with torch.cuda.stream(self.s_head):
                disp, stats = F_.disp_head_forward(cost_lr, self.maxdisp, want_stats=want_stats)
        for t in (x, y):
            t.record_stream(self.s_cv)
        cost_lr.record_stream(self.s_head)
        # the outputs live in the side streams' allocator pools but are consumed on the caller's stream (after
        # join()): tell the caching allocator, so a freed block is not recycled under a running consumer
        for t in (cost, disp, stats):
            if t is not None:
                t.record_stream(cur)
        return cost, disp, stats

    def join(self):
        """Make the current stream wait for everything enqueued so far."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.s_cv)
        cur.wait_stream(self.s_head)


class HostPipeline:
    def __init__(self, maxdisp: int = 192, device: torch.device | str = "cuda", depth: int = 2):
        self.device = torch.device(device)
        self.maxdisp = maxdisp
        self.cv = CostVolume(maxdisp)
        self.head = Disp(maxdisp)
        self.depth = depth
        self._copy = torch.cuda.Stream(self.device)
        self._d2h = torch.cuda.Stream(self.device)
        self._slots = None

    def _ensure(self, x_h, cl_h):
        key = (tuple(x_h.shape), tuple(cl_h.shape))
        if self._slots is not None and self._key == key:
            return
        self._key = key
        b, _, dl, hl, wl = cl_h.shape
        self._slots = []
        for _ in range(self.depth):
            self._slots.append({
                "x": torch.empty(x_h.shape, dtype=torch.float32, device=self.device),
                "y": torch.empty(x_h.shape, dtype=torch.float32, device=self.device),
                "cl": torch.empty(cl_h.shape, dtype=torch.float32, device=self.device),
                "disp_h": torch.empty((b, 3 * hl, 3 * wl), dtype=torch.float32).pin_memory(),
                "ready": torch.cuda.Event(), "free": torch.cuda.Event(), "done": torch.cuda.Event(),
            })
        self._n = 0

    def submit(self, x_h: torch.Tensor, y_h: torch.Tensor, cost_lr_h: torch.Tensor, keep_volume: bool = False):
        """Enqueue one step.  Inputs are (pinned) HOST fp32 tensors: left/right features
        [B,C,Hf,Wf] and the matching cost [B,1,Dl,Hl,Wl].  Returns the slot whose ``disp_h``
        (pinned host [B,3Hl,3Wl]) is valid after ``slot['done'].synchronize()``."""
        self._ensure(x_h, cost_lr_h)
        slot = self._slots[self._n % self.depth]
        self._n += 1
        compute = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._copy):
            self._copy.wait_event(slot["free"])    # previous user of this slot's device buffers finished
            self._copy.wait_event(slot["done"])    # ... and its host result was copied out
            slot["x"].copy_(x_h, non_blocking=True)
            slot["y"].copy_(y_h, non_blocking=True)
            slot["cl"].copy_(cost_lr_h, non_blocking=True)
            slot["ready"].record(self._copy)
        compute.wait_event(slot["ready"])
        with torch.no_grad():
            cost = self.cv(slot["x"], slot["y"])       # [B,2C,Df,Hf,Wf], stays on the device (Matching Net input)
            disp = self.head(slot["cl"])
        slot["free"].record(compute)
        self._d2h.wait_event(slot["free"])
        with torch.cuda.stream(self._d2h):
            slot["disp_h"].copy_(disp, non_blocking=True)
            disp.record_stream(self._d2h)
            slot["done"].record(self._d2h)
        if keep_volume:
            slot["cost"] = cost
        return slot

    def run(self, x_h, y_h, cost_lr_h) -> torch.Tensor:
        """Synchronous convenience call: returns the pinned host disparity map."""
        slot = self.submit(x_h, y_h, cost_lr_h)
        slot["done"].synchronize()
        return slot["disp_h"]

    def drain(self):
        for s in self._slots or []:
            s["done"].synchronize()
