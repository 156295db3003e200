#!/usr/bin/env python
"""One backprop-through-4-steps of Conv3 at 64 x 257^2 (BASELINE config 4 shape per GPU) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file out.csv python tools/bptt_profile.py"""
import sys, importlib, argparse, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
npde = importlib.import_module("neural-pde-solver_b200")
opt = argparse.Namespace(iterator="conv", activation="clamp", conv_n_layers=3, mg_n_layers=3, mg_pre_smoothing=2,
                         mg_post_smoothing=2, cg_n_iters=4, geometry="square", is_train=True, optimizer="adam",
                         lr_init=1e-3, max_iter_steps=20, max_iter_steps_from_gt=20, lambda_gt=0.0)
B, N = 64, 257
rng = np.random.RandomState(0)
x = torch.from_numpy(rng.rand(B, N, N).astype(np.float32)).cuda()
bc = torch.from_numpy(rng.rand(B, 4).astype(np.float32)).cuda()
x = npde.utils.set_boundary(x, bc)
gt = x.clone()
m = npde.HeatModel(opt, full_bptt=True)
m.iterator.cuda()
it = m.iterator
y = it.unrolled(x, bc, None, 4)
torch.nn.MSELoss()(y, gt).backward()
torch.cuda.synchronize()
