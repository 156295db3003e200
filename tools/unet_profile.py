#!/usr/bin/env python
"""A few U-Net(3,2,2) steps at 1 x 1025^2 (the published protocol's shape) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file out.csv python tools/unet_profile.py"""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
npde = importlib.import_module("neural-pde-solver_b200")
B, N = int(os.environ.get("UP_B", 1)), int(os.environ.get("UP_N", 1025))
rng = np.random.RandomState(0)
x = torch.from_numpy(rng.rand(B, N, N).astype(np.float32)).cuda()
bc = torch.from_numpy(rng.rand(B, 4).astype(np.float32)).cuda()
x = npde.utils.set_boundary(x, bc)
it = npde.UNetIterator("clamp", int(os.environ.get("UP_L", 3)), 2, 2).cuda()
with torch.no_grad():
    for p in it.parameters():
        p.copy_(torch.from_numpy(((rng.rand(1, 1, 3, 3) - 0.5) * 0.1).astype(np.float32)))
    y, _, _ = it.iterate(x, bc, None, 3)
torch.cuda.synchronize()
