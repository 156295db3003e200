#!/usr/bin/env python
"""The reference's whole workflow on one B200, end to end through this package:

  1. generation.py   -> a square training set (default 65^2) and square test sets (default 257/513/1025^2),
                        solved to fd_error < 1e-5 on the device;
  2. train.py        -> a UNet(3,2,2) iterator trained as scripts/train.sh does (batch 32, Adam 1e-3, up to 20
                        unrolled iterations, random initialisation), data read back through data.get_data_loader;
  3. runtime.py      -> the published-number protocol of scripts/runtime.sh (avg initialisation, 800 evaluation
                        steps, 1 % relative error): `num` problems one at a time (B = 1) and as one batch.

The figures it prints are the B200 counterparts of visualizations/plot_curves.py:29-46 ("time for 100 solves");
the authors' CPU numbers are quoted next to them.  One JSON line per grid size.
"""
import argparse
import copy
import importlib
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
npde = importlib.import_module("neural-pde-solver_b200")

PUBLISHED_OURS_S = {257: 13.757, 513: 66.134, 1025: 299.918}     # visualizations/plot_curves.py:31 (CPU, unspecified)
PUBLISHED_FENICS_S = {257: 17.170, 513: 74.683, 1025: 307.176}   # visualizations/plot_curves.py:30


def quiet(fn, *a, **k):
    """The reference's drivers print per-iteration progress; keep the tool's output to its JSON lines."""
    saved = sys.stdout
    sys.stdout = open(os.devnull, "w")
    try:
        return fn(*a, **k)
    finally:
        sys.stdout.close()
        sys.stdout = saved


def generate(root, image_size, n_runs, device, batch_size=16):
    opt = argparse.Namespace(save_dir=root, batch_size=batch_size, save_every=1, n_frames=1, n_runs=n_runs,
                             image_size=image_size, max_temp=100, poisson=0, geometry="square", use_model=1,
                             device=str(device), init_ckpt="")
    t0 = time.time()
    iters = quiet(npde.generation.generate_square, opt)
    return opt.save_dir, iters, time.time() - t0


def train(dset_path, args, device):
    opt = argparse.Namespace(iterator="unet", activation="clamp", conv_n_layers=3, mg_n_layers=3, mg_pre_smoothing=2,
                             mg_post_smoothing=2, cg_n_iters=4, geometry="square", is_train=True, optimizer="adam",
                             lr_init=1e-3, max_iter_steps=20, max_iter_steps_from_gt=0, lambda_gt=0.0,
                             dset_path=dset_path, max_temp=100, data_limit=0, poisson=0, batch_size=32, n_workers=0,
                             initialization="random")
    model = npde.HeatModel(opt)
    model.iterator.to(device)
    loader = npde.data.get_data_loader(opt)
    step, losses, t0 = 0, [], time.time()
    for epoch in range(args.epochs):
        model.setup(is_train=True)
        for batch in loader:
            x = npde.utils.initialize(batch["x"], batch["bc"], opt.initialization)
            out = model.train(x, batch["final"], batch["bc"], None)
            losses.append(out["loss"]["loss_x"])
            step += 1
            if args.max_train_steps and step >= args.max_train_steps:
                break
        if args.max_train_steps and step >= args.max_train_steps:
            break
    model.setup(is_train=False)
    return opt, model, {"train_steps": step, "train_s": round(time.time() - t0, 2),
                        "loss_first": float(np.mean(losses[:10])), "loss_last": float(np.mean(losses[-10:]))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="+", default=[257, 513, 1025])
    ap.add_argument("--num", type=int, default=100)
    ap.add_argument("--train_size", type=int, default=65)
    ap.add_argument("--train_runs", type=int, default=40)
    ap.add_argument("--epochs", type=int, default=20)
    ap.add_argument("--max_train_steps", type=int, default=0)
    ap.add_argument("--n_evaluation_steps", type=int, default=800)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--seed", type=int, default=666)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    device = torch.device(args.device)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    root = tempfile.mkdtemp(prefix="npde_protocol_")
    lines = []
    try:
        train_dir, iters, gen_s = generate(os.path.join(root, "train"), args.train_size, args.train_runs, device)
        model_opt, model, info = train(train_dir, args, device)
        info.update({"stage": "train", "train_set": "%d x 16 problems at %d^2" % (args.train_runs, args.train_size),
                     "generation_s": round(gen_s, 2), "jacobi_iterations_per_run": iters[:4]})
        lines.append(info)
        print(json.dumps(info), flush=True)

        run_opt = copy.deepcopy(model_opt)
        run_opt.is_train = False
        run_opt.initialization = "avg"                  # scripts/runtime.sh:6
        run_opt.n_evaluation_steps = args.n_evaluation_steps
        for n in args.sizes:
            n_runs = args.num // 16 + 1                     # runtime.py:14
            dset, iters, gen_s = generate(os.path.join(root, "test%d" % n), n, n_runs, device)
            data = npde.runtime.get_data("square", args.num, dset, 100)
            times, steps = quiet(npde.runtime.runtime, run_opt, model, data, device=device)
            t_batch, first = npde.runtime.runtime_batched(run_opt, model, data, device=device)
            line = {"stage": "runtime", "grid": n, "problems": int(data["bc"].size(0)), "solved": len(times),
                    "steps_mean": float(np.mean(steps)) if steps else None,
                    "steps_max": int(max(steps)) if steps else None,
                    "sum_s": float(sum(times)),
                    "adjusted_time_s": float(np.mean(times) * args.num) if times else None,
                    "batched_time_s": float(t_batch), "batched_solved": int((first >= 0).sum()),
                    "published_ours_cpu_s": PUBLISHED_OURS_S.get(n), "published_fenics_cpu_s": PUBLISHED_FENICS_S.get(n),
                    "generation_s": round(gen_s, 2), "jacobi_iterations_per_run": iters,
                    "protocol": "runtime.py:41-85, threshold 1 %% rel. RMS, init avg, <= %d steps, "
                                "B=1 timed with CUDA events around one fused iterate(k=steps)" % args.n_evaluation_steps}
            lines.append(line)
            print(json.dumps(line), flush=True)
    finally:
        shutil.rmtree(root, ignore_errors=True)
    if args.out:
        with open(args.out, "w") as fh:
            for line in lines:
                fh.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
