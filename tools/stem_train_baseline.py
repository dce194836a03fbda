"""Baseline for the next round's training-mode fused stem: the reference's layer sequence (volume built by our
cost-volume kernel -> nn.Conv3d(24, 12, 3, pad 1, bias=False) -> BatchNorm3d(train) -> ReLU) forward + backward
through cuDNN / PyTorch on this GPU, TF32 on and off.   python tools/stem_train_baseline.py [B Hf Wf maxdisp]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rag_b200.modules import CostVolume  # noqa: E402

b, hf, wf, md = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (4, 96, 192, 192)
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(b, 12, hf, wf, device=dev, generator=g, requires_grad=True)
y = torch.randn(b, 12, hf, wf, device=dev, generator=g, requires_grad=True)
conv = torch.nn.Conv3d(24, 12, 3, 1, 1, bias=False).to(dev)
bn = torch.nn.BatchNorm3d(12).to(dev).train()
cv = CostVolume(md)
gout = torch.randn(b, 12, md // 3, hf, wf, device=dev, generator=g)


def step():
    out = torch.relu(bn(conv(cv(x, y))))
    out.backward(gout)


for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10 if tf32 else 3
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"what": "volume + Conv3d(24->12) + BatchNorm3d(train) + ReLU, forward + backward (cuDNN/PyTorch)",
                      "tf32": tf32, "B": b, "Hf": hf, "Wf": wf, "maxdisp": md, "ms_per_step": round(e0.elapsed_time(e1) / n, 3),
                      "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}))
