"""Time rag_conv3d_c1_fwd (last_3_3d forward) and check it against an fp64 convolution on a sub-block.
Usage: python tools/conv_time.py"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_b200 import _cabi
L = _cabi.lib()
for (b, c, d, h, w) in [(8, 12, 64, 160, 320), (4, 12, 64, 96, 192), (4, 12, 64, 128, 416), (1, 12, 61, 50, 76)]:
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(b, c, d, h, w, device="cuda", generator=g)
    wt = torch.randn(1, c, 3, 3, 3, device="cuda", generator=g) * 0.1
    out = torch.empty(b, 1, d, h, w, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    run = lambda: _cabi.check(L.rag_conv3d_c1_fwd(x.data_ptr(), wt.data_ptr(), out.data_ptr(), b, c, d, h, w, st), "fwd")
    for _ in range(3): run()
    ms = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        ms.append(round(e0.elapsed_time(e1) / 20, 4))
    ref = torch.nn.functional.conv3d(x[:1].double(), wt.double(), padding=1)
    err = float((out[:1].double() - ref).abs().max() / ref.abs().max())
    print(json.dumps({"shape": [b, c, d, h, w], "ms": ms, "max_norm_err_vs_fp64": err}))
