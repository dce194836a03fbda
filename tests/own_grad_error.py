"""Own error of the head-backward variants against the fp64 evaluator (max-norm relative), per sigma.
usage (GPU box): python tests/own_grad_error.py [variants...]   -- test infrastructure: imports oracle/."""
import os, sys
import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import rag_oracle as O  # noqa: E402
from rag_b200 import functional as F_  # noqa: E402

variants = [int(v) for v in sys.argv[1:]] or [0, 1, 2, 3]
for (dl, hl, wl, md) in ((64, 24, 64, 192), (96, 12, 40, 288)):
    for sigma in (1.0, 5.0, 20.0):
        g = torch.Generator().manual_seed(7)
        cost = torch.randn((1, 1, dl, hl, wl), generator=g) * sigma
        gd = torch.randn((1, 3 * hl, 3 * wl), generator=g) * (torch.rand((1, 3 * hl, 3 * wl), generator=g) < 0.3)
        _, g64 = O.disp_head_grad_f64(cost[:, 0].numpy(), gd.numpy(), md)
        _, gref = O.disp_head_grad_ref(cost, gd, md)
        ref_err = np.abs(gref.numpy()[:, 0] - g64).max() / np.abs(g64).max()
        disp, stats = F_.disp_head_forward(cost.cuda(), md, want_stats=True)
        row = []
        for v in variants:
            out = F_.disp_head_backward(cost.cuda(), gd.cuda(), disp, stats, md, variant=v).cpu().numpy()[:, 0]
            row.append(f"v{v} {np.abs(out - g64).max() / np.abs(g64).max():.2e}")
        print(f"Dl={dl} md={md} sigma={sigma}: reference fp32 {ref_err:.2e} | " + "  ".join(row))
