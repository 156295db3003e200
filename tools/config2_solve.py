#!/usr/bin/env python
"""BASELINE configs[1]: the learned multigrid-shaped iterator (UNetIterator) on 257 x 257 L-shape and cylinder
geometries, batch 256, iterated to fd_error < 1e-6, next to Jacobi and Multigrid from the same start.

    python tools/config2_solve.py [--batch 256] [--grid 257] > profiles/r2_config2.jsonl

Geometries: the package's port of utils/geometries.py (seeded numpy draw order of the reference).  Weights:
tests/golden/trained_unet322.npz -- a UNet(3,2,2) TRAINED BY THE REFERENCE's own HeatModel.train on 33^2 square
problems (tests/golden/make_golden.py gen_trained), i.e. the paper's "train small, run on other sizes and
geometries" protocol.  Start: zero interior + G (heat_model.py:104-106).  One JSON line per (geometry, iterator):
iterations (multiple of the check interval), seconds (CUDA-synchronised wall clock), converged, per-problem first
crossing statistics."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
npde = importlib.import_module("neural-pde-solver_b200")


def trained_unet(dev):
    with np.load(os.path.join(ROOT, "tests", "golden", "trained_unet322.npz")) as z:
        g = {k: z[k] for k in z.files}
    _, levels, pre, post = (int(v) for v in g["meta"])
    it = npde.UNetIterator("clamp", levels, pre, post).to(dev)
    it.load_state_dict({k[2:].replace("/", "."): torch.from_numpy(v) for k, v in g.items() if k.startswith("w/")})
    return it.to(dev), (levels, pre, post)


def solve(it, x, bc, tol, every, cap, per_problem):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        res = npde.utils.solve_to_tolerance(it, x, bc, None, tol=tol, check_every=every, max_iters=cap,
                                            per_problem=per_problem)
    torch.cuda.synchronize()
    return res, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--grid", type=int, default=257)
    ap.add_argument("--tol", type=float, default=1e-6)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    B, N = args.batch, args.grid
    unet, shape = trained_unet(dev)
    for geometry in ("Lshape", "cylinders"):
        np.random.seed(666)
        _, v, m = npde.utils.get_geometry(geometry, N, B, 1)
        bc = torch.from_numpy(np.stack([v, m], 1).astype(np.float32)).to(dev)
        x0 = npde.utils.initialize(torch.zeros(B, N, N, device=dev), bc, "zero")
        mg = npde.MultigridIterator(4, 8, 8)
        for name, it, every, cap in (("unet(%d,%d,%d) trained by the reference at 33^2" % shape, unet, 10, 20000),
                                     ("multigrid(4,8,8)", mg, 10, 5000),
                                     ("jacobi", npde.JacobiIterator(), 2000, 400000)):
            for itr in (it,):
                itr.is_bc_mask = True
            res, dt = solve(it, x0.clone(), bc, args.tol, every, cap, per_problem=False)
            line = {"config": "256 x 257^2 %s, zero start, fd_error < %g" % (geometry, args.tol), "iterator": name,
                    "batch": B, "grid": N, "iterations": int(res["iterations"]), "seconds": dt,
                    "converged": bool(res["converged"]), "final_residual": float(res["history"][-1][1]),
                    "check_every": every,
                    "cell_updates_per_s": B * N * N * float(res["iterations"]) / dt}
            if name.startswith("unet") and res["converged"]:
                resp, _ = solve(it, x0.clone(), bc, args.tol, every, cap, per_problem=True)
                fb = resp["first_below"].numpy()
                line["first_crossing"] = {"min": int(fb.min()), "median": float(np.median(fb)), "max": int(fb.max())}
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
