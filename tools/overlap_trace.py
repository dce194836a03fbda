"""Timeline of one cost-volume launch and one head launch on two streams (do they really co-run?)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200 import _cabi
L = _cabi.lib()
b, c, hf, wf, df, md = 8, 12, 160, 320, 64, 192
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(b, c, hf, wf, device="cuda", generator=g); y = torch.randn(b, c, hf, wf, device="cuda", generator=g)
cl = torch.randn(b, 1, df, hf, wf, device="cuda", generator=g)
cost = torch.empty(b, 2 * c, df, hf, wf, device="cuda")
disp = torch.empty(b, 3 * hf, 3 * wf, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
E = lambda: torch.cuda.Event(enable_timing=True)

def cv(v, st): L.rag_cost_volume_fwd_v(x.data_ptr(), y.data_ptr(), cost.data_ptr(), b, c, df, hf, wf, v, st.cuda_stream)
def hd(v, st): L.rag_disp_head_fwd_v(cl.data_ptr(), disp.data_ptr(), None, b, df, hf, wf, md, v, st.cuda_stream)

def trace(cvv, hv, head_first, n_head=1):
    for _ in range(2):
        cv(cvv, s1); hd(hv, s2)
    torch.cuda.synchronize()
    e0, a0, a1, b0, b1 = E(), E(), E(), E(), E()
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    if head_first:
        b0.record(s2)
        for _ in range(n_head): hd(hv, s2)
        b1.record(s2)
        a0.record(s1); cv(cvv, s1); a1.record(s1)
    else:
        a0.record(s1); cv(cvv, s1); a1.record(s1)
        b0.record(s2)
        for _ in range(n_head): hd(hv, s2)
        b1.record(s2)
    torch.cuda.synchronize()
    return {"cv": cvv, "head": hv, "head_first": head_first, "n_head": n_head,
            "cv_start": round(e0.elapsed_time(a0), 3), "cv_end": round(e0.elapsed_time(a1), 3),
            "head_start": round(e0.elapsed_time(b0), 3), "head_end": round(e0.elapsed_time(b1), 3)}

for cvv in (28, 29):
    for hv in (10, 14):
        for hf_ in (False, True):
            print(json.dumps(trace(cvv, hv, hf_)), flush=True)
print(json.dumps(trace(29, 14, False, n_head=2)))
