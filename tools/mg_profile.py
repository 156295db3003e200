#!/usr/bin/env python
"""A few V-cycles for an ncu launch list:  MG_B=8 MG_N=4097 ncu --metrics gpu__time_duration.sum ... python tools/mg_profile.py"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
npde = importlib.import_module("neural-pde-solver_b200")
B, N = int(os.environ.get("MG_B", 8)), int(os.environ.get("MG_N", 4097))
rng = np.random.RandomState(0)
x = torch.from_numpy(rng.rand(B, N, N).astype(np.float32)).cuda()
bc = torch.from_numpy(rng.rand(B, 4).astype(np.float32)).cuda()
x = npde.utils.set_boundary(x, bc)
mg = npde.MultigridIterator(int(os.environ.get("MG_L", 6)), 8, 8)
with torch.no_grad():
    for _ in range(3):
        x = mg(x, bc, None)
torch.cuda.synchronize()
