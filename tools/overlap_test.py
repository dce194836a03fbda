"""Does the (HBM-bound) cost-volume kernel overlap with the (SFU/issue-bound) head kernel on two streams?"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200 import functional as F_
b, c, hf, wf, df, md = 8, 12, 160, 320, 64, 192
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(b, c, hf, wf, device="cuda", generator=g); y = torch.randn(b, c, hf, wf, device="cuda", generator=g)
cl = torch.randn(b, 1, df, hf, wf, device="cuda", generator=g)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
cost = torch.empty(b, 2 * c, df, hf, wf, device="cuda")
from rag_b200 import _cabi
L = _cabi.lib()
disp = torch.empty(b, 3 * hf, 3 * wf, device="cuda")

def run(cv_variant, head_first, n=20, overlap=True, head_variant=-1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(n):
        def cv():
            L.rag_cost_volume_fwd_v(x.data_ptr(), y.data_ptr(), cost.data_ptr(), b, c, df, hf, wf, cv_variant, (s1 if overlap else torch.cuda.current_stream()).cuda_stream)
        def hd():
            L.rag_disp_head_fwd_v(cl.data_ptr(), disp.data_ptr(), None, b, df, hf, wf, md, head_variant, (s2 if overlap else torch.cuda.current_stream()).cuda_stream)
        if head_first: hd(); cv()
        else: cv(); hd()
    if overlap:
        ev1, ev2 = torch.cuda.Event(), torch.cuda.Event()
        ev1.record(s1); ev2.record(s2)
        torch.cuda.current_stream().wait_event(ev1); torch.cuda.current_stream().wait_event(ev2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for cvv in (32,):
    for hv in (9, 10, 11, 12):
        run(cvv, False, head_variant=hv); run(cvv, True, head_variant=hv)
        print(json.dumps({"cv_variant": cvv, "head_variant": hv, "serial_ms": round(run(cvv, False, overlap=False, head_variant=hv), 4),
                          "overlap_cv_first_ms": round(run(cvv, False, head_variant=hv), 4),
                          "overlap_head_first_ms": round(run(cvv, True, head_variant=hv), 4)}), flush=True)
