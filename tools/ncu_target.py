#!/usr/bin/env python
"""DEVELOPMENT: a short, single-process run of the Conv iterator for ncu captures.
    python tools/ncu_target.py [B N L K]      (default 64 1025 3 12)"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
npde = importlib.import_module("neural-pde-solver_b200")
B, N, L, K = (int(v) for v in (sys.argv[1:5] + ["64", "1025", "3", "12"][len(sys.argv) - 1:]))
rng = np.random.RandomState(0)
bc = torch.from_numpy(rng.rand(B, 4).astype(np.float32)).cuda()
x = npde.utils.set_boundary(torch.from_numpy(rng.rand(B, N, N).astype(np.float32)).cuda(), bc)
it = npde.ConvIterator("clamp", L).cuda()
with torch.no_grad():
    for l in it.layers:
        l.weight.copy_(torch.from_numpy(((rng.rand(3, 3) - 0.5) * 0.2).astype(np.float32)).view(1, 1, 3, 3))
    out = torch.empty_like(x)
    for _ in range(2):
        it.iterate(x, bc, None, K, out=out)
torch.cuda.synchronize()
print("ok", float(out.double().sum()))
