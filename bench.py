#!/usr/bin/env python
"""Contract benchmark of the RAG stereo hot path (cost volume + disparity regression) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload infer|train] [--impl reference]

A step = one pass of the hot path over one batch of synthetic stereo pairs.  Default workload
(BASELINE.json configs[1]): inference batch 8 per GPU at DrivingStereo half-res 400x880, which the
reference pads to 480x960 at eval (dataloaders/stereo_dataset.py:95-96; its Feature Net cannot run
400x880) -> features [8,12,160,320], volume [8,24,64,160,320], head input [8,1,64,160,320],
maxdisp 192.  `--workload train` is configs[2]: fwd+bwd, batch 4 per GPU at 288x576.

Prints ONE JSON line (rank 0).  Inference workloads are timed on the two-stream schedule of
rag_b200.pipeline.OverlappedPath (the cost volume of a batch shares the SMs with the disparity head of
another): `value` = pairs/s of K such steps with inputs resident in HBM (CUDA events,
max over ranks); `path.serial` = the same K steps with the kernels back to back on one stream, which is also
where the per-kernel durations and `roofline` (the dominant kernel, cost-volume forward, against the measured
HBM peak) come from.  The training workload (fwd+bwd) is timed serially.  `e2e` = the same metric through
rag_b200.pipeline.HostPipeline with pinned HOST buffers (H2D of the inputs and D2H of the disparity inside the
timed region); `cpu_baseline` = the oracle port of the reference timed on this box's host cores; `next_rows`
= the SURVEY.md section 8f rows (fused stem, last_3_3d conv) beside the headline.  `--impl reference` times the
reference's CPU path (oracle port: the reference is pure PyTorch, so the port IS its op sequence) on the
same config.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "stereo pairs/s (cost vol + disp regression)"
WORKLOADS = {
    # name: (B per GPU, C, Hf, Wf, Df, maxdisp, backward?, description)
    "infer": (8, 12, 160, 320, 64, 192, False, "inference batch 8/GPU at 400x880 padded to 480x960 (reference eval pad), maxdisp 192"),
    "train": (4, 12, 96, 192, 64, 192, True, "training step fwd+bwd batch 4/GPU at 288x576 crop, maxdisp 192"),
    # configs[3]: 32 pairs routed over 4 grown scene paths = 4 sub-batches of 8 per step (the hot path is path-independent)
    "router": (32, 12, 160, 320, 64, 192, False, "scene-router inference, 32 pairs/GPU at 480x960 as 4 path sub-batches of 8, maxdisp 192"),
    # configs[4]: resolution / disparity sweep
    "sweep192": (4, 12, 128, 416, 64, 192, False, "inference batch 4/GPU at 384x1248, maxdisp 192"),
    "sweep288": (4, 12, 128, 416, 96, 288, False, "inference batch 4/GPU at 384x1248, maxdisp 288"),
}


def alg_bytes(c, hf, wf, df):
    """Algorithmic bytes per pair."""
    cv = 4 * (2 * c * hf * wf + 2 * c * df * hf * wf)
    head_f = 4 * (df * hf * wf + 9 * hf * wf)
    head_b = 4 * (2 * df * hf * wf + 9 * hf * wf)
    return cv, head_f, head_b


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram read+write bytes per launch of the dominant kernel from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


# --------------------------------------------------------------------------------------------
# CPU reference arm (oracle port)
# --------------------------------------------------------------------------------------------
def cpu_reference(workload, budget_s=12.0, min_reps=2):
    import torch

    from oracle import rag_oracle as O

    b, c, hf, wf, df, md, bwd, _ = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(1, c, hf, wf, generator=g)
    y = torch.randn(1, c, hf, wf, generator=g)
    cl = torch.randn(1, 1, df, hf, wf, generator=g)

    def step():
        if not bwd:
            with torch.no_grad():
                O.cost_volume_ref(x, y, md)
                O.disp_head_ref(cl, md)
        else:
            xr, yr, cr = x.clone().requires_grad_(True), y.clone().requires_grad_(True), cl.clone().requires_grad_(True)
            cost = O.cost_volume_ref(xr, yr, md)
            cost.backward(torch.ones_like(cost))
            d = O.disp_head_ref(cr, md)
            d.backward(torch.ones_like(d))

    step()  # warm-up
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < min_reps or (time.perf_counter() < t_end and len(times) < 50):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    best = min(times)
    return {
        "value": 1.0 / best, "unit": "pairs/s", "cores": cores, "kind": "port",
        "sample": f"1 pair of the {workload} workload ({3*hf}x{3*wf}, maxdisp {md}, {'fwd+bwd' if bwd else 'fwd'}), best of {len(times)} after 1 warm-up, torch CPU {torch.get_num_threads()} threads; "
                  "pairs are independent so pairs/s does not depend on batch",
    }, times


def torch_cuda_reference(workload, dev, reps=5):
    """The reference's own op sequence (oracle port) run by PyTorch eager ON THE SAME GPU: the honest GPU
    baseline.  One pair per call (its temporaries are ~1.8 GB/pair at 480x960)."""
    import torch

    from oracle import rag_oracle as O

    b, c, hf, wf, df, md, bwd, _ = WORKLOADS[workload]
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(1, c, hf, wf, generator=g).to(dev)
    y = torch.randn(1, c, hf, wf, generator=g).to(dev)
    cl = torch.randn(1, 1, df, hf, wf, generator=g).to(dev)

    def step():
        if not bwd:
            with torch.no_grad():
                O.cost_volume_ref(x, y, md)
                O.disp_head_ref(cl, md)
        else:
            xr, yr, cr = x.clone().requires_grad_(True), y.clone().requires_grad_(True), cl.clone().requires_grad_(True)
            cost = O.cost_volume_ref(xr, yr, md)
            cost.backward(torch.ones_like(cost))
            d = O.disp_head_ref(cr, md)
            d.backward(torch.ones_like(d))

    step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    torch.cuda.empty_cache()
    return {"value": 1e3 / ms, "unit": "pairs/s", "ms_per_pair": ms,
            "what": "reference op sequence (rag_model.py:375-383,18-44) through PyTorch eager CUDA kernels on this GPU, 1 pair per call"}


def fused_stem_row(x, y, df, md, sub):
    """SURVEY.md section 8f rank 1 (reported beside the headline, not part of `value`): cost volume + first
    Matching-Net layer (ConvBR_3d 2C->C, 3x3x3, eval-mode BN, ReLU).  Fused kernel (volume never
    materialised, fp32) vs the reference composition on the materialised volume through cuDNN (TF32 default)."""
    import torch

    from rag_b200 import functional as F_
    from rag_b200.fused_stem import cv_stem_forward

    dev = x.device
    xs, ys = x[:sub], y[:sub]
    c = x.shape[1]
    g = torch.Generator(device=dev).manual_seed(11)
    conv = torch.nn.Conv3d(2 * c, c, 3, padding=1, bias=False).to(dev)
    bn = torch.nn.BatchNorm3d(c).to(dev).eval()
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.3, generator=g)
        bn.running_var.uniform_(0.5, 2.0, generator=g)
        scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
        shift = bn.bias - bn.running_mean * scale

        def t(fn, n):
            fn()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) / n

        fused = t(lambda: cv_stem_forward(xs, ys, conv.weight, scale, shift, True, md), 10)
        ref = t(lambda: torch.relu_(bn(conv(F_.cost_volume_forward(xs, ys, df)))), 3)
        a = cv_stem_forward(xs[:1], ys[:1], conv.weight, scale, shift, True, md)
        old = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        b_ = torch.relu_(bn(conv(F_.cost_volume_forward(xs[:1], ys[:1], df))))
        torch.backends.cudnn.allow_tf32 = old
        err = ((a - b_).abs().max() / b_.abs().max()).item()
        # rank 2: last_3_3d (Conv3d C -> 1, 3x3x3), the producer of the head's input
        from rag_b200.last_conv import conv3d_c1_forward
        feat = torch.randn(sub, c, df, x.shape[2], x.shape[3], device=dev, generator=g)
        last = torch.nn.Conv3d(c, 1, 3, padding=1, bias=False).to(dev)
        lc = t(lambda: conv3d_c1_forward(feat, last.weight), 10)
        lc_ref = t(lambda: last(feat), 3)
        torch.backends.cudnn.allow_tf32 = False
        lc_err = ((conv3d_c1_forward(feat[:1], last.weight) - last(feat[:1])).abs().max() / last(feat[:1]).abs().max()).item()
        torch.backends.cudnn.allow_tf32 = old
        del feat
    torch.cuda.empty_cache()
    return {"stem3d0": {"what": "cost volume + stem3d0 (Conv3d 24->12 3x3x3 + BN(eval) + ReLU), %d pairs" % sub,
                        "fused_ms": round(fused, 4), "cost_volume_plus_cudnn_tf32_ms": round(ref, 4), "speedup": round(ref / fused, 1),
                        "max_rel_err_vs_fp32_cudnn": err},
            "last_3_3d": {"what": "last_3_3d (Conv3d %d->1 3x3x3, fp32 direct convolution), %d pairs" % (c, sub),
                          "ms": round(lc, 4), "cudnn_tf32_ms": round(lc_ref, 4), "speedup": round(lc_ref / lc, 1),
                          "max_rel_err_vs_fp32_cudnn": lc_err}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    b, c, hf, wf, df, md, bwd, desc = WORKLOADS[args.workload]
    # each "step" = one pair on the host cores; run warmup+steps of them, bounded
    base, _ = cpu_reference(args.workload, budget_s=0.0, min_reps=max(1, min(args.steps, 8)))
    v = base["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "desc": desc, "note": "reference CPU path = oracle port of rag_model.py:375-383,18-44 (pure PyTorch ops), 1 pair per step on all host cores"},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# clocks sampler (NVML, in-process: the timed region is far shorter than nvidia-smi's period)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml as nv

            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from rag_b200 import _cabi
    from rag_b200 import functional as F_
    from rag_b200.modules import CostVolume, Disp
    from rag_b200.pipeline import HostPipeline, OverlappedPath

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    b, c, hf, wf, df, md, bwd, desc = WORKLOADS[args.workload]
    K, Wm = args.steps, max(args.warmup, 3)
    # identical bits for CPU reference and GPU: generate on the host with a fixed seed, then copy
    g = torch.Generator().manual_seed(1234 + rank)
    x_h = torch.randn(b, c, hf, wf, generator=g).pin_memory()
    y_h = torch.randn(b, c, hf, wf, generator=g).pin_memory()
    cl_h = torch.randn(b, 1, df, hf, wf, generator=g).pin_memory()
    x, y, cl = x_h.to(dev), y_h.to(dev), cl_h.to(dev)
    cv_mod, head_mod = CostVolume(md), Disp(md)
    if bwd:
        gcost = torch.randn(b, 2 * c, df, hf, wf, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
        gd_h = torch.randn(b, 3 * hf, 3 * wf, generator=g) * (torch.rand(b, 3 * hf, 3 * wf, generator=g) < 0.3)
        gdisp = gd_h.to(dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    marks = [[ev() for _ in range(5)] for _ in range(K)]

    sub = 8 if args.workload == "router" else b   # router: one call per scene path

    def step(mk=None):
        if mk: mk[0].record()
        if not bwd:
            cost = [F_.cost_volume_forward(x[i:i + sub], y[i:i + sub], df) for i in range(0, b, sub)]
            if mk: mk[1].record()
            disp = [F_.disp_head_forward(cl[i:i + sub], md, want_stats=False)[0] for i in range(0, b, sub)]
            if mk: mk[2].record()
            return cost, disp
        cost = F_.cost_volume_forward(x, y, df)
        if mk: mk[1].record()
        disp, stats = F_.disp_head_forward(cl, md, want_stats=True)
        if mk: mk[2].record()
        gcl = F_.disp_head_backward(cl, gdisp, disp, stats, md)
        if mk: mk[3].record()
        gx, gy = F_.cost_volume_backward(gcost, c)
        if mk: mk[4].record()
        return cost, disp, gcl, gx, gy

    for _ in range(Wm):
        out = step()
    del out

    # ---- serial pass: the kernels back to back on one stream, each bracketed by events.  This is where the
    # per-kernel durations (roofline) come from; for the training workload it is also the timed region. ----
    def serial_pass():
        barrier()
        t0, t1 = ev(), ev()
        t0.record()
        for k in range(K):
            out = step(marks[k])
        t1.record()
        torch.cuda.synchronize()
        del out
        return t0.elapsed_time(t1)

    # ---- overlapped pass (inference workloads): cost volume and disparity head on two streams, the
    # schedule of rag_b200.pipeline.OverlappedPath for a stream of batches ----
    opath = OverlappedPath(md, dev) if not bwd else None

    def overlapped_step():
        outs = [opath.step(x[i:i + sub], y[i:i + sub], cl[i:i + sub]) for i in range(0, b, sub)]
        return outs

    def overlapped_pass():
        for _ in range(Wm):
            outs = overlapped_step()
        opath.join()
        barrier()
        t0, t1 = ev(), ev()
        t0.record()
        for k in range(K):
            outs = overlapped_step()
        opath.join()
        t1.record()
        torch.cuda.synchronize()
        del outs
        return t0.elapsed_time(t1)

    sampler = ClockSampler(local) if rank == 0 else None
    serial_ms = None
    if not bwd:
        serial_ms = serial_pass()            # untimed for `value`: kernel durations only
    barrier()
    n0 = _cabi.launch_count()
    if sampler: sampler.start()
    total_ms = overlapped_pass() if not bwd else serial_pass()
    if sampler: clocks = sampler.stop()
    launches = _cabi.launch_count() - n0 - (0 if bwd else Wm * 2 * (b // sub))   # minus the warm-up launches of the pass
    barrier()
    if world > 1:
        t = torch.tensor([total_ms, serial_ms if serial_ms is not None else total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t[0].item())
        serial_ms = float(t[1].item()) if serial_ms is not None else None
    value = world * b * K / (total_ms * 1e-3)

    # per-kernel average durations (serial pass: each kernel alone on the GPU, back to back)
    n_seg = 4 if bwd else 2
    seg_ms = [sum(marks[k][i].elapsed_time(marks[k][i + 1]) for k in range(K)) / K for i in range(n_seg)]
    cvb, hfb, hbb = alg_bytes(c, hf, wf, df)
    peak, peak_src = measured_peak()
    n_sub = b // sub
    cv_ms = seg_ms[0] / n_sub                       # average duration of ONE cost-volume launch
    achieved = cvb * sub / (cv_ms * 1e-3) / 1e9
    lean = wf % 4 == 0
    roofline = {
        "bound": "hbm", "kernel": ("cv_fwd_lean_kernel<256,2,true>" if lean else "cv_fwd_kernel") + " (cost-volume forward)",
        "achieved": round(achieved, 1), "peak": peak,
        "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": ncu_traffic("cv_fwd_lean_kernel" if lean else "cv_fwd_kernel"),
        "peak_source": peak_src, "alg_bytes_per_launch": cvb * sub, "avg_launch_ms": round(cv_ms, 5),
        "timed_in": "serial pass of %d steps in this run (kernels back to back on one stream, CUDA events around each)" % K,
        "note": "a write-only stream: the measured peak is a COPY (read+write) figure, a pure fill of the same bytes runs at ~7.45 TB/s",
    }
    path_bytes = (cvb + hfb + ((cvb + hbb) if bwd else 0)) * b
    kernels = {"cv_fwd_ms": seg_ms[0], "head_fwd_ms": seg_ms[1]}
    if bwd:
        kernels.update({"head_bwd_ms": seg_ms[2], "cv_bwd_ms": seg_ms[3]})
    step_ms = total_ms / K
    path = {
        "alg_bytes_per_step": path_bytes, "achieved_GBps": round(path_bytes / (step_ms * 1e-3) / 1e9, 1),
        "frac_of_hbm_peak": round(path_bytes / (step_ms * 1e-3) / 1e9 / peak, 4),
        "schedule": ("serial: the four kernels back to back on one stream" if bwd else
                     "overlapped: cost volume and disparity head on two streams (rag_b200.pipeline.OverlappedPath); the two kernels of a step have no "
                     "data dependence (in the network the Matching Net sits between them: head of batch i runs beside the volume of batch i+1)"),
        "kernel_ms": {k: round(v, 5) for k, v in kernels.items()},
        "note": "the head kernels are FP32-pipe bound (one exp2 + 7-12 FP32 ops per pixel per low-res bin), not HBM bound; see DESIGN.md",
    }
    if serial_ms is not None:
        path["serial"] = {"ms_per_step": round(serial_ms / K, 5), "pairs_per_s": round(world * b * K / (serial_ms * 1e-3), 1),
                          "frac_of_hbm_peak": round(path_bytes / (serial_ms / K * 1e-3) / 1e9 / peak, 4)}

    # ---- e2e: HOST buffers through the public pipeline (H2D + kernels + D2H per step) ----
    pipe = HostPipeline(md, dev)
    Ke = max(3, min(K, 50))
    for _ in range(3):
        pipe.submit(x_h, y_h, cl_h)
    pipe.drain()
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    last = None
    for _ in range(Ke):
        last = pipe.submit(x_h, y_h, cl_h)
    pipe.drain()
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1)
    barrier()
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    checksum = float(last["disp_h"].double().mean())
    h2d = (x_h.numel() + y_h.numel() + cl_h.numel()) * 4
    d2h = last["disp_h"].numel() * 4
    e2e = {"value": world * b * Ke / (e2e_ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": Ke, "ms_per_step": e2e_ms / Ke, "mode": "forward (inference) through rag_b200.pipeline.HostPipeline, pinned host buffers, double-buffered copies",
           "mean_disp": checksum}

    cpu = None
    torch_gpu = None
    next_rows = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_reference(args.workload, budget_s=12.0)
        torch_gpu = torch_cuda_reference(args.workload, dev)
        if not bwd:
            next_rows = fused_stem_row(x, y, df, md, sub)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "desc": desc, "pairs_per_gpu": b, "features": [b, c, hf, wf],
                       "volume": [b, 2 * c, df, hf, wf], "head_in": [b, 1, df, hf, wf], "maxdisp": md,
                       "l2": "no explicit flush: every step streams %.2f GB (>> 126 MB L2) through HBM" % (path_bytes / 1e9),
                       "sharding": "stereo pairs across ranks, no data-path collective"},
            "roofline": roofline, "path": path, "cpu_baseline": cpu, "torch_cuda_reference": torch_gpu, "next_rows": next_rows, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="infer")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
