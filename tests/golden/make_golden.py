"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports ``/root/reference/src`` (models.rag_model.Network / Disp / DisparityRegression and
automl.mdenas_basicmodel), applies the one monkeypatch needed to execute it on CPU
(``torch.cuda.current_device = lambda: 'cpu'`` -- rag_model.py:26 builds its arange on
"the current CUDA device"), runs the reference's own ``forward`` code on small seeded
inputs and stores inputs + outputs (+ autograd gradients) as ``.npz`` under tests/golden/.

The cost volume is captured from the reference's own ``Network.forward`` loop by
handing it a stub ``feature``/``matching`` (identity feature, matching records its input)
so lines rag_model.py:375-383 themselves run -- not a restatement of them.

The reference does not exist on the GPU box, so nothing at test time imports it: tests
read only the committed ``.npz`` files.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.path.insert(0, REF)
    torch.cuda.current_device = lambda: "cpu"  # rag_model.py:26
    import models.rag_model as rm  # noqa: E402
    import utilstool.metrics as metrics  # noqa: E402

    return rm, metrics


def _ref_cost_volume(rm, x, y, maxdisp):
    """Run the reference's Network.forward (rag_model.py:369-387) with stubbed
    feature/matching/disp so only lines 375-383 do work; returns the volume it built."""
    holder = {}
    stub = types.SimpleNamespace()
    stub.maxdisp = maxdisp
    feats = iter([x, y])
    stub.feature = lambda img, task_arch, path: next(feats)

    def matching(cost, task_arch, path):
        holder["cost"] = cost
        return cost

    stub.matching = matching
    stub.disp = lambda c: c
    rm.Network.forward(stub, None, None, 0)
    return holder["cost"]


def main():
    rm, metrics = _import_reference()
    g = torch.Generator().manual_seed(1234)

    # ---- cost volume + gradient -----------------------------------------------------
    cases_cv = {
        # name: (B, C, Hf, Wf, maxdisp)
        "cv_b2_c12_h6_w20_md24": (2, 12, 6, 20, 24),
        "cv_b1_c12_h4_w70_md192": (1, 12, 4, 70, 192),  # Wf > Df, Wf % 4 == 2
        "cv_b1_c3_h3_w10_md48": (1, 3, 3, 10, 48),  # Wf < Df (empty slices)
        "cv_b1_c12_h2_w37_md30": (1, 12, 2, 37, 30),  # odd Wf
    }
    for name, (b, c, hf, wf, md) in cases_cv.items():
        x = torch.randn(b, c, hf, wf, generator=g)
        y = torch.randn(b, c, hf, wf, generator=g)
        xr = x.clone().requires_grad_(True)
        yr = y.clone().requires_grad_(True)
        cost = _ref_cost_volume(rm, xr, yr, md)
        # wide-dynamic-range upstream gradient exposes the accumulation order
        gc = torch.randn(cost.shape, generator=g) * torch.exp(3 * torch.randn(cost.shape, generator=g))
        cost.backward(gc)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            x=x.numpy(), y=y.numpy(), maxdisp=md, cost=cost.detach().numpy(),
            gcost=gc.numpy(), gx=xr.grad.numpy(), gy=yr.grad.numpy(),
        )
        print(name, tuple(cost.shape))

    # ---- head + gradient ------------------------------------------------------------
    cases_head = {
        # name: (B, Dl, Hl, Wl, maxdisp, sigma)
        "head_b2_d8_h4_w6_md24_s1": (2, 8, 4, 6, 24, 1.0),
        "head_b1_d64_h5_w7_md192_s1": (1, 64, 5, 7, 192, 1.0),
        "head_b1_d64_h4_w5_md192_s5": (1, 64, 4, 5, 192, 5.0),
        "head_b1_d48_h3_w4_md192_s1": (1, 48, 3, 4, 192, 1.0),  # depth != maxdisp/3
        "head_b1_d96_h2_w3_md288_s1": (1, 96, 2, 3, 288, 1.0),
    }
    for name, (b, dl, hl, wl, md, sigma) in cases_head.items():
        cost = (sigma * torch.randn(b, 1, dl, hl, wl, generator=g)).requires_grad_(True)
        head = rm.Disp(md)
        disp = head(cost)
        gd = torch.randn(disp.shape, generator=g) * (torch.rand(disp.shape, generator=g) < 0.3).float()
        disp.backward(gd)
        # DisparityRegression alone (rag_model.py:18-29) on a random probability volume
        p = torch.softmax(torch.randn(b, md, 3 * hl, 3 * wl, generator=g), dim=1).contiguous()
        reg = rm.DisparityRegression(md)(p)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            cost=cost.detach().numpy(), maxdisp=md, disp=disp.detach().numpy(),
            gdisp=gd.numpy(), gcost=cost.grad.numpy(), p=p.numpy(), reg=reg.numpy(),
        )
        print(name, tuple(disp.shape))

    # ---- loss + metrics (approaches/rag.py:418-430, utilstool/metrics.py) -----------------
    import torch.nn.functional as F
    import warnings

    warnings.filterwarnings("ignore")
    b, h, w, md = 3, 12, 20, 192
    gt = torch.rand(b, h, w, generator=g) * 260 - 30  # some <=0, some >= maxdisp
    gt[2] = torch.where(torch.rand(h, w, generator=g) < 0.97, torch.full((h, w), 250.0), gt[2])  # image 2 gets skipped
    est = gt + 4 * torch.randn(b, h, w, generator=g)
    mask = (gt < md) & (gt > 0)
    import contextlib, io

    with contextlib.redirect_stdout(io.StringIO()):
        out = {
            "loss": F.smooth_l1_loss(est[mask], gt[mask], size_average=True).item(),
            "EPE": metrics.EPE_metric(est, gt, mask).item(),
            "D1": metrics.D1_metric(est, gt, mask).item(),
            "Thres1": metrics.Thres_metric(est, gt, mask, 1.0).item(),
            "Thres2": metrics.Thres_metric(est, gt, mask, 2.0).item(),
            "Thres3": metrics.Thres_metric(est, gt, mask, 3.0).item(),
        }
    np.savez_compressed(os.path.join(HERE, "metrics_b3_h12_w20.npz"), est=est.numpy(), gt=gt.numpy(), maxdisp=md,
                        **{k: np.float64(v) for k, v in out.items()})
    print("metrics", out)

    # ---- eval-time staging (data_io.py:6-13 + stereo_dataset.py:88-102) -------------------
    from dataloaders.data_io import get_transform
    from PIL import Image

    rng = np.random.RandomState(7)
    img = rng.randint(0, 256, size=(10, 14, 3), dtype=np.uint8)
    t = get_transform()(Image.fromarray(img)).numpy()
    top_pad, right_pad = 12 - 10, 18 - 14
    padded = np.pad(t, ((0, 0), (top_pad, 0), (0, right_pad)), mode="constant", constant_values=0)  # np.lib.pad == np.pad
    np.savez_compressed(os.path.join(HERE, "stage_h10_w14.npz"), img=img, top_pad=top_pad, right_pad=right_pad, out=padded)
    print("stage", padded.shape)

    # ================= added later: own generator, so that the files above stay bit-identical =================
    g2 = torch.Generator().manual_seed(4321)

    # ---- head cases the x3 tiled kernels take (maxdisp == 3*Dl, Wl % 4 == 0) -------------------------------
    cases_head2 = {
        "head_b1_d64_h6_w8_md192_s1": (1, 64, 6, 8, 192, 1.0),
        "head_b2_d32_h5_w12_md96_s5": (2, 32, 5, 12, 96, 5.0),
    }
    for name, (b, dl, hl, wl, md, sigma) in cases_head2.items():
        cost = (sigma * torch.randn(b, 1, dl, hl, wl, generator=g2)).requires_grad_(True)
        disp = rm.Disp(md)(cost)
        gd = torch.randn(disp.shape, generator=g2) * (torch.rand(disp.shape, generator=g2) < 0.3).float()
        disp.backward(gd)
        p = torch.softmax(torch.randn(b, md, 3 * hl, 3 * wl, generator=g2), dim=1).contiguous()
        reg = rm.DisparityRegression(md)(p)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            cost=cost.detach().numpy(), maxdisp=md, disp=disp.detach().numpy(),
            gdisp=gd.numpy(), gcost=cost.grad.numpy(), p=p.numpy(), reg=reg.numpy(),
        )
        print(name, tuple(disp.shape))

    # ---- next rows: the reference's own ConvBR_3d layers ----------------------
    # stem3d0 = ConvBR_3d(2C, C, 3, 1, 1) on the volume built by Network.forward (rag_model.py:234,341,375-383) in
    # eval mode, and last_3_3d = ConvBR_3d(C, 1, 3, 1, 1, bn=False, relu=False) (rag_model.py:269).
    from automl.operations_3d import ConvBR_3d

    torch.manual_seed(99)
    c, hf, wf, md = 12, 5, 12, 24
    x = torch.randn(1, c, hf, wf, generator=g2)
    y = torch.randn(1, c, hf, wf, generator=g2)
    stem = ConvBR_3d(2 * c, c, 3, 1, 1).eval()
    with torch.no_grad():
        stem.bn.running_mean.copy_(0.3 * torch.randn(c, generator=g2))
        stem.bn.running_var.copy_(0.5 + torch.rand(c, generator=g2))
        stem.bn.weight.copy_(0.5 + torch.rand(c, generator=g2))
        stem.bn.bias.copy_(0.2 * torch.randn(c, generator=g2))
        vol = _ref_cost_volume(rm, x, y, md)
        out = stem(vol)
    np.savez_compressed(
        os.path.join(HERE, "stem_b1_c12_h5_w12_md24.npz"), x=x.numpy(), y=y.numpy(), maxdisp=md,
        weight=stem.conv.weight.detach().numpy(), bn_weight=stem.bn.weight.detach().numpy(), bn_bias=stem.bn.bias.detach().numpy(),
        bn_mean=stem.bn.running_mean.numpy(), bn_var=stem.bn.running_var.numpy(), bn_eps=stem.bn.eps, out=out.numpy(),
    )
    print("stem", tuple(out.shape))
    last = ConvBR_3d(c, 1, 3, 1, 1, bn=False, relu=False).eval()
    feat = torch.randn(1, c, 18, 9, 12, generator=g2)
    with torch.no_grad():
        mat = last(feat)
    np.savez_compressed(os.path.join(HERE, "last3_b1_c12_d18_h9_w12.npz"), feat=feat.numpy(),
                        weight=last.conv.weight.detach().numpy(), out=mat.numpy())
    print("last_3_3d", tuple(mat.shape))


if __name__ == "__main__":
    main()
