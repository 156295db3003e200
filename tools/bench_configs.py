#!/usr/bin/env python
"""Secondary measurements (not the bench.py headline line): the other BASELINE.json configurations and
the building-block kernels, device resident, CUDA events, best of 3.  One JSON line per measurement.

    python tools/bench_configs.py [--quick] > gpurun_out/configs.jsonl
"""
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
DEV = torch.device("cuda:0")


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def square_problem(B, N, seed=0, poisson=False):
    rng = np.random.RandomState(seed)
    bc = torch.from_numpy(rng.rand(B, 4).astype(np.float32)).to(DEV)
    x = torch.rand(B, N, N, device=DEV)
    x = npde.utils.set_boundary(x, bc)
    f = None
    if poisson:
        f = torch.zeros(B, N, N, device=DEV)
        f[:, 1:-1, 1:-1] = -torch.rand(B, N - 2, N - 2, device=DEV) * 3.0 / (N * N)
    return x, bc, f


def mask_problem(B, N, kind, seed=0):
    """L-shape / cylinders: bc = [values, mask] (utils/geometries.py semantics: mask 1 on Dirichlet cells)."""
    rng = np.random.RandomState(seed)
    m = np.zeros((B, N, N), np.float32)
    m[:, 0, :] = m[:, -1, :] = m[:, :, 0] = m[:, :, -1] = 1
    v = np.zeros((B, N, N), np.float32)
    ring = rng.rand(B).astype(np.float32)
    if kind == "Lshape":
        h = N // 2
        m[:, :h, h:] = 1
        v[m > 0] = np.repeat(ring, (m > 0).reshape(B, -1).sum(1))
    else:
        yy, xx = np.mgrid[0:N, 0:N]
        for b in range(B):
            v[b][m[b] > 0] = ring[b]
            for _ in range(3):
                cy, cx, r = rng.randint(N // 5, 4 * N // 5, 2).tolist() + [rng.randint(N // 16, N // 8)]
                disk = (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
                m[b][disk] = 1
                v[b][disk] = rng.rand()
    bc = torch.from_numpy(np.stack([v, m], 1)).to(DEV)
    x = npde.utils.initialize(torch.zeros(B, N, N, device=DEV), bc, "zero")
    return x, bc


def conv_iterator(L, act="clamp", scale=0.2, seed=1):
    it = npde.ConvIterator(act, L).to(DEV)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for l in it.layers:
            l.weight.copy_(((torch.rand(1, 1, 3, 3, generator=g) - 0.5) * scale).to(DEV))
    return it


def unet_iterator(L, pre, post, act="clamp", scale=0.1, seed=2):
    it = npde.UNetIterator(act, L, pre, post).to(DEV)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in it.parameters():
            p.copy_(((torch.rand(1, 1, 3, 3, generator=g) - 0.5) * scale).to(DEV))
    return it


def emit(name, ms, cell_updates, **extra):
    rec = {"name": name, "ms": ms, "G_cell_updates_per_s": cell_updates / ms / 1e6}
    rec.update(extra)
    print(json.dumps(rec), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    torch.manual_seed(0)
    with torch.no_grad():
        # ---- config 3 family: 64 x 1025^2 -------------------------------------------------------
        B, N, K = 64, 1025, 50
        x, bc, f = square_problem(B, N, poisson=True)
        out = torch.empty_like(x)
        for L in (1, 2, 3, 4):
            it = conv_iterator(L)
            ms = timed(lambda: it.iterate(x, bc, None, K, out=out))
            emit("conv%d_1025_b64_iterate" % L, ms / K, B * N * N, bytes_per_cell=8)
        it3 = conv_iterator(3)
        ms = timed(lambda: it3.iterate(x, bc, f, K, out=out))
        emit("conv3_1025_b64_poisson_iterate", ms / K, B * N * N, bytes_per_cell=12)
        ms = timed(lambda: it3(x, bc, None))
        emit("conv3_1025_b64_single_step_reference_layout", ms, B * N * N, bytes_per_cell=8)
        ms = timed(lambda: it3.iterate(x, bc, None, K, want_fd=True, out=out))
        emit("conv3_1025_b64_iterate_with_fd_rows", ms / K, B * N * N)
        jac = npde.JacobiIterator()
        for depth in (1, 2, 4, 8):
            ms = timed(lambda: jac.iterate(x, bc, None, 64, out=out, temporal_depth=depth))
            emit("jacobi_1025_b64_depth%d" % depth, ms / 64, B * N * N, bytes_per_cell=8.0 / depth)
        for depth in (1, 4):  # with a Poisson source: 12 B per cell-update at depth 1
            ms = timed(lambda: jac.iterate(x, bc, f, 64, out=out, temporal_depth=depth))
            emit("jacobi_1025_b64_poisson_depth%d" % depth, ms / 64, B * N * N, bytes_per_cell=12.0 / depth)
        ms = timed(lambda: npde.utils.fd_error(x, bc, None))
        emit("fd_error_1025_b64", ms, B * N * N, bytes_per_cell=4)
        mg = npde.MultigridIterator(6, 8, 8)
        ms = timed(lambda: mg(x, bc, None))
        emit("multigrid(6,8,8)_vcycle_1025_b64", ms, B * N * N, smoothing_updates_per_cell=(8 + 8) * 4 / 3)
        del x, f, out

        # ---- config 1: 65^2, batch 32, 800 evaluation steps -------------------------------------
        B, N = 32, 65
        x, bc, _ = square_problem(B, N)
        it = conv_iterator(3)
        ms = timed(lambda: it.iterate(x, bc, None, 800))
        emit("conv3_65_b32_800_steps", ms, B * N * N * 800, steps_per_s=800 / ms * 1e3)
        gt = x.clone()
        ms = timed(lambda: it.iterate(x, bc, None, 800, gt=gt, want_l2=True))
        emit("conv3_65_b32_800_steps_with_l2_rows", ms, B * N * N * 800)
        ms = timed(lambda: jac.iterate(x, bc, None, 800))
        emit("jacobi_65_b32_800_steps", ms, B * N * N * 800)

        # ---- config 2: 257^2 mask geometries, batch 256 ------------------------------------------
        B, N = (64 if args.quick else 256), 257
        for kind in ("Lshape", "cylinders"):
            x, bc = mask_problem(B, N, kind)
            un = unet_iterator(3, 2, 2)
            un.is_bc_mask = True
            ms = timed(lambda: un(x, bc, None))
            emit("unet(3,2,2)_%s_257_b%d_step" % (kind, B), ms, B * N * N)
            mgm = npde.MultigridIterator(4, 8, 8)
            mgm.is_bc_mask = True
            ms = timed(lambda: mgm(x, bc, None))
            emit("multigrid(4,8,8)_%s_257_b%d_vcycle" % (kind, B), ms, B * N * N)
            cv = conv_iterator(3)
            cv.is_bc_mask = True
            ms = timed(lambda: cv(x, bc, None))
            emit("conv3_%s_257_b%d_step" % (kind, B), ms, B * N * N, bytes_per_cell=16)
            ms = timed(lambda: cv.iterate(x, bc, None, 40))
            emit("conv3_%s_257_b%d_iterate" % (kind, B), ms / 40, B * N * N, bytes_per_cell=16)
            jm = npde.JacobiIterator()
            jm.is_bc_mask = True
            ms = timed(lambda: jm.iterate(x, bc, None, 20))
            emit("jacobi_%s_257_b%d" % (kind, B), ms / 20, B * N * N, bytes_per_cell=16)
        # time to 1e-6 residual, Jacobi vs multigrid warm start + Jacobi is the reference protocol; here: Jacobi, square
        B, N = 64, 257
        x, bc, _ = square_problem(B, N)
        t0 = time.perf_counter()
        x = npde.utils.initialize(x, bc, "avg")  # the reference's evaluation start (test.py:115); random noise can stall fp32 Jacobi
        res = npde.utils.solve_to_tolerance(npde.JacobiIterator(), x, bc, None, tol=1e-6, check_every=1000,
                                            max_iters=400000)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"name": "jacobi_257_b64_time_to_1e-6_residual", "seconds": dt, "iterations": res["iterations"],
                          "converged": res["converged"], "final_residual": res["history"][-1][1]}), flush=True)

    if not args.quick:
        # BASELINE's second metric at the headline size: Jacobi to a 1e-6 residual on 64 x 1025^2 (temporally blocked,
        # one scalar read back per 2000 iterations).  The CPU reference needs ~0.77 s per iteration.
        B, N = 64, 1025
        x, bc, _ = square_problem(B, N)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x = npde.utils.initialize(x, bc, "avg")
        res = npde.utils.solve_to_tolerance(npde.JacobiIterator(), x, bc, None, tol=1e-6, check_every=2000,
                                            max_iters=2000000)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"name": "jacobi_1025_b64_time_to_1e-6_residual", "seconds": dt, "iterations": res["iterations"],
                          "converged": res["converged"], "final_residual": res["history"][-1][1],
                          "G_cell_updates_per_s": B * N * N * res["iterations"] / dt / 1e9}), flush=True)

    # ---- config 4: training step, 257^2, batch 64 per GPU, K = 20 --------------------------------
    import argparse as _ap
    opt = _ap.Namespace(iterator="conv", activation="clamp", conv_n_layers=3, mg_n_layers=3, mg_pre_smoothing=2,
                        mg_post_smoothing=2, cg_n_iters=4, geometry="square", is_train=True, optimizer="adam",
                        lr_init=1e-3, max_iter_steps=20, max_iter_steps_from_gt=20, lambda_gt=0.0)
    B, N = 64, 257
    x, bc, _ = square_problem(B, N)
    gt = x.clone()
    model = npde.HeatModel(opt)
    model.iterator.to(DEV)
    np.random.seed(0)
    model.train(x, gt, bc, None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    steps = 0
    np.random.seed(1)
    for _ in range(n):
        model.train(x, gt, bc, None)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"name": "train_step_conv3_257_b64_K20 (N~U{1..20} fused no-grad steps + 1 differentiable step, Adam)",
                      "ms_per_train_step": dt / n * 1e3}), flush=True)
    # BASELINE config 4 proper: gradient through ALL N unrolled iterations (N ~ U{1..20}, mean 10.5)
    modelf = npde.HeatModel(opt, full_bptt=True)
    modelf.iterator.to(DEV)
    modelf.train(x, gt, bc, None)
    torch.cuda.synchronize()
    np.random.seed(1)
    t0 = time.perf_counter()
    for _ in range(n):
        modelf.train(x, gt, bc, None)
    torch.cuda.synchronize()
    print(json.dumps({"name": "train_step_conv3_257_b64_K20_full_bptt (gradient through all N~U{1..20} iterations, Adam)",
                      "ms_per_train_step": (time.perf_counter() - t0) / n * 1e3}), flush=True)
    opt.iterator = "unet"
    modelu = npde.HeatModel(opt)
    modelu.iterator.to(DEV)
    modelu.train(x, gt, bc, None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        modelu.train(x, gt, bc, None)
    torch.cuda.synchronize()
    print(json.dumps({"name": "train_step_unet(3,2,2)_257_b64_K20", "ms_per_train_step": (time.perf_counter() - t0) / n * 1e3}),
          flush=True)

    # ---- config 5: 4097^2 ------------------------------------------------------------------------
    with torch.no_grad():
        B, N = (2 if args.quick else 8), 4097
        x, bc, _ = square_problem(B, N)
        mg = npde.MultigridIterator(6, 8, 8)
        ms = timed(lambda: mg(x, bc, None), reps=2)
        emit("multigrid(6,8,8)_vcycle_4097_b%d" % B, ms, B * N * N)
        it3 = conv_iterator(3)
        out = torch.empty_like(x)
        ms = timed(lambda: it3.iterate(x, bc, None, 10, out=out), reps=2)
        emit("conv3_4097_b%d_iterate" % B, ms / 10, B * N * N, bytes_per_cell=8)
        un = unet_iterator(6, 2, 2)
        ms = timed(lambda: un(x, bc, None), reps=2)
        emit("unet(6,2,2)_4097_b%d_step" % B, ms, B * N * N)


if __name__ == "__main__":
    main()
