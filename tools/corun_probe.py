"""Which resource of a co-running kernel slows the HBM-bound cost-volume kernel?  Synthetic co-runners (FMA / LDS /
MUFU spin at 12 warps per SM) vs paced global reads.  Result: profiles/r1_overlap_schedule.md section 6."""
import ctypes, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200 import _cabi
L = _cabi.lib()
Fl = ctypes.CDLL(os.path.join(ROOT, "tools", "libcorun_probe.so"))   # nvcc -O3 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC tools/corun_probe.cu -o tools/libcorun_probe.so
Fl.launch_corun.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p]
Fl.launch_reader.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p]
big = torch.randn(64 * 1024 * 1024, device="cuda")   # 256 MB
b, c, hf, wf, df = 8, 12, 160, 320, 64
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(b, c, hf, wf, device="cuda", generator=g); y = torch.randn(b, c, hf, wf, device="cuda", generator=g)
cost = torch.empty(b, 2 * c, df, hf, wf, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
E = lambda: torch.cuda.Event(enable_timing=True)
def cv(v, st): L.rag_cost_volume_fwd_v(x.data_ptr(), y.data_ptr(), cost.data_ptr(), b, c, df, hf, wf, v, st.cuda_stream)
for mb in (26, 105, 210, 420):
    for _ in range(2):
        torch.cuda.synchronize()
        e0, a0, a1 = E(), E(), E()
        e0.record(); s1.wait_event(e0); s2.wait_event(e0)
        a0.record(s1); cv(32, s1); a1.record(s1)
        ctas = 444
        iters = mb * 1024 * 1024 // (ctas * 128 * 16)
        gap = int(0.38e-3 * 1.965e9 / iters)
        Fl.launch_reader(big.data_ptr(), ctas, iters, gap, s2.cuda_stream)
        torch.cuda.synchronize()
    print(json.dumps({"cv": 32, "corunner": "paced global reads", "MB_read_during_cv": mb, "cv_ms": round(a0.elapsed_time(a1), 4)}), flush=True)
for cvv in (32,):
    for _ in range(3): cv(cvv, s1)
    torch.cuda.synchronize()
    for mode, name in ((-1, "alone"), (0, "FMA spin"), (1, "LDS spin"), (2, "MUFU spin"), (3, "FMA+LDS")):
        for ctas in (3, 2, 1):
            if mode < 0 and ctas != 3: continue
            torch.cuda.synchronize()
            e0, a0, a1 = E(), E(), E()
            e0.record(); s1.wait_event(e0); s2.wait_event(e0)
            a0.record(s1); cv(cvv, s1); a1.record(s1)                 # cv first (persistent, resident at once)
            if mode >= 0: Fl.launch_corun(mode, 148 * ctas, 58 * 1024, 1200000, s2.cuda_stream)   # ~0.6 ms of spinning
            torch.cuda.synchronize()
            print(json.dumps({"cv": cvv, "corunner": name, "ctas_per_sm": ctas if mode >= 0 else 0, "cv_ms": round(a0.elapsed_time(a1), 4)}), flush=True)
