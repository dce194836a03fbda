"""How long does the reference's first Matching-Net layer (stem3d0 = Conv3d 24->12, 3x3x3, pad 1, + BN + ReLU)
take on the materialised volume?  (cuDNN through PyTorch, with and without TF32.)"""
import json, sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200 import functional as F_
b, c, hf, wf, df = 8, 12, 160, 320, 64
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(b, c, hf, wf, device="cuda", generator=g); y = torch.randn(b, c, hf, wf, device="cuda", generator=g)
conv = torch.nn.Conv3d(24, 12, 3, padding=1, bias=False).cuda()
bn = torch.nn.BatchNorm3d(12).cuda().eval()
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
with torch.no_grad():
    cost = F_.cost_volume_forward(x, y, df)
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        for cl in (False, True):
            cc = cost.contiguous(memory_format=torch.channels_last_3d) if cl else cost
            print(json.dumps({"tf32": tf32, "channels_last": cl, "conv3d_ms": round(t(lambda: conv(cc)), 3),
                              "conv_bn_relu_ms": round(t(lambda: torch.relu_(bn(conv(cc)))), 3)}))
    print(json.dumps({"cost_volume_ms": round(t(lambda: F_.cost_volume_forward(x, y, df)), 3)}))
