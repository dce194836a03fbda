"""Timing of the Matching-Net tail rows (upsample_6 fwd+bwd, last_3_3d fwd+bwd) at B=4 288x576.  python tools/prof_tail.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import tail_rows  # noqa: E402

print(tail_rows(torch.device("cuda:0")))
