#!/usr/bin/env python
"""DEVELOPMENT: time of one backward step of the Conv iterator (npde_conv_step_bwd) and of backprop through k unrolled
steps, fused streaming kernel vs the layered kernels (NPDE_BWD_LAYERED=1).
    python tools/ab_bwd.py [B N L K]      (default 64 257 3 20)"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
npde = importlib.import_module("neural-pde-solver_b200")
B, N, L, K = (int(v) for v in (sys.argv[1:5] + ["64", "257", "3", "20"][len(sys.argv) - 1:]))
rng = np.random.RandomState(0)
bc = torch.from_numpy(rng.rand(B, 4).astype(np.float32)).cuda()
x = npde.utils.set_boundary(torch.from_numpy(rng.rand(B, N, N).astype(np.float32)).cuda(), bc)
gout = torch.from_numpy(rng.rand(B, N, N).astype(np.float32) - 0.5).cuda()
it = npde.ConvIterator("clamp", L).cuda()
with torch.no_grad():
    for l in it.layers:
        l.weight.copy_(torch.from_numpy(((rng.rand(3, 3) - 0.5) * 0.2).astype(np.float32)).view(1, 1, 3, 3))


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def one_step():
    xg = x.clone().requires_grad_(True)
    y = it(xg, bc, None)
    y.backward(gout)
    return xg.grad


def unrolled():
    xg = x.clone().requires_grad_(True)
    y = it.unrolled(xg, bc, None, K)
    y.backward(gout)
    return xg.grad


for mode in ("stream", "layered"):
    if mode == "layered":
        os.environ["NPDE_BWD_LAYERED"] = "1"
    else:
        os.environ.pop("NPDE_BWD_LAYERED", None)
    for l in it.layers:
        l.weight.grad = None
    g1 = one_step()
    gw1 = [l.weight.grad.clone() for l in it.layers]
    t1 = timed(one_step)
    tk = timed(unrolled)
    with torch.no_grad():
        tf = timed(lambda: it.iterate(x, bc, None, K, temporal_depth=1))
    print("%-8s B=%d N=%d L=%d: fwd+bwd of one step %.1f us; %d unrolled steps fwd+bwd %.2f ms (forward-only %d steps %.2f ms "
          "-> backward ~%.1f us per step)   |gx| %.6g  gw[0][0,0] %.6g" % (
              mode, B, N, L, t1 * 1e3, K, tk, K, tf, (tk - tf) / K * 1e3, float(g1.abs().sum()), float(gw1[0].flatten()[0])))
