"""One fused training-stem step (B=4, 288x576) for ncu / per-kernel timing.  python tools/prof_stem_train.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import stem_train_row  # noqa: E402

print(stem_train_row(torch.device("cuda:0")))
