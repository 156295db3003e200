#!/usr/bin/env python
"""Data-parallel training over NCCL on real GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_train_check.py

Every rank trains on its slice of a (64 x world) x 257^2 batch -- BASELINE config 4 at world = 8: batch 512,
K = 20 -- with HeatModel(opt, world=True) (weight-gradient all-reduce over NVLink); rank 0 also trains a
single-process model on the full batch.  Checks: replicas stay bit-identical, and equal the full-batch run to
1e-6; prints the time per training step.  NPDE_FULL_BPTT=1: gradient through all unrolled iterations."""
import argparse
import importlib
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
npde = importlib.import_module("neural-pde-solver_b200")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    opt = argparse.Namespace(iterator="conv", activation="clamp", conv_n_layers=3, mg_n_layers=3, mg_pre_smoothing=2,
                             mg_post_smoothing=2, cg_n_iters=4, geometry="square", is_train=True, optimizer="adam",
                             lr_init=1e-3, max_iter_steps=20, max_iter_steps_from_gt=20, lambda_gt=0.0)
    full_bptt = os.environ.get("NPDE_FULL_BPTT", "0") == "1"
    B, N = 64 * world, 257
    rng = np.random.RandomState(0)
    bc = rng.rand(B, 4).astype(np.float32)
    x = rng.rand(B, N, N).astype(np.float32)
    gt = (x * 0.5 + 0.25).astype(np.float32)
    w0 = ((rng.rand(3, 1, 1, 3, 3) - 0.5) * 0.2).astype(np.float32)

    def make(world_flag):
        m = npde.HeatModel(opt, world=world_flag, full_bptt=full_bptt)
        m.iterator.to(dev)
        with torch.no_grad():
            for l, w in zip(m.iterator.layers, w0):
                l.weight.copy_(torch.from_numpy(w))
        return m

    per = B // world
    sl = slice(rank * per, (rank + 1) * per)
    xs, gts, bcs = (torch.from_numpy(a[sl]).to(dev) for a in (x, gt, bc))
    model = make(True)
    np.random.seed(7)  # identical unroll lengths on every rank
    model.train(xs, gts, bcs, None)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    steps = 10
    for _ in range(steps):
        out = model.train(xs, gts, bcs, None)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    w = torch.stack([l.weight.detach().reshape(9) for l in model.iterator.layers])
    gathered = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(gathered, w)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        ref = make(None)
        np.random.seed(7)
        for _ in range(steps + 1):
            ref.train(torch.from_numpy(x).to(dev), torch.from_numpy(gt).to(dev), torch.from_numpy(bc).to(dev), None)
        wr = torch.stack([l.weight.detach().reshape(9) for l in ref.iterator.layers])
        print("world %d, global batch %d, full_bptt %s: replicas identical: %s; max |w_dp - w_single| = %.3g; "
              "%.2f ms per DP train step; loss %.6g"
              % (world, B, full_bptt, same, float((w - wr).abs().max()), dt * 1e3, out["loss"]["loss_x"]))
        assert same and float((w - wr).abs().max()) < 1e-6
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
