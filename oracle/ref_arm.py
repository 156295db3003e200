#!/usr/bin/env python
"""TEST / MEASUREMENT INFRASTRUCTURE: times the UNMODIFIED reference's own CPU path for the bench workload.

    CUDA_VISIBLE_DEVICES='' python oracle/ref_arm.py --B 64 --N 1025 --layers 3 --act clamp --warmup 1 --steps 3

oracle/_ref/ holds a verbatim copy of the reference's `utils/`, `models/` and `args/` packages, made by
`__graft_entry__.build()` from /root/reference in the build container (git-ignored build output, shipped to the
GPU box with the snapshot like the built .so files; never part of the repository's history).  This script imports
`models.iterators.ConvIterator` from there -- with the two viz-only modules that are not installed stubbed
(matplotlib, tensorboardX: SURVEY.md 8c) -- and times `ConvIterator.forward` (models/iterators/conv_iterator.py:33-51)
under torch.no_grad() on the host cores, one Phi_H application of the whole batch per step.  Prints one JSON object.
Run in a subprocess with CUDA_VISIBLE_DEVICES='' so that utils/heat_utils.py:24-27 keeps its kernels on the CPU.
"""
import argparse
import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF, "models", "iterators", "conv_iterator.py")) and \
        os.path.exists(os.path.join(REF, "utils", "heat_utils.py"))


def import_reference():
    os.environ.setdefault("HOME", "/root")  # utils/build.py, args/base_args.py read it at import time
    for name in ("matplotlib", "matplotlib.pyplot", "tensorboardX"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["tensorboardX"].SummaryWriter = object
    sys.path.insert(0, REF)
    import utils  # noqa: F401  (the reference's package)
    from models.iterators import ConvIterator
    return ConvIterator


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--N", type=int, default=1025)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--act", default="clamp")
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--seed", type=int, default=666)
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    import numpy as np
    import torch
    assert not torch.cuda.is_available(), "run with CUDA_VISIBLE_DEVICES='' (the reference moves its kernels to the GPU)"
    if args.threads > 0:
        torch.set_num_threads(args.threads)
    ConvIterator = import_reference()
    sys.path.insert(0, os.path.dirname(HERE))
    import bench  # the same synthetic problem as the GPU arm (make_problem)
    x, bc, w = bench.make_problem(args.B, args.N, args.seed)
    it = ConvIterator(args.act, args.layers)
    with torch.no_grad():
        for layer, wl in zip(it.layers, w):
            layer.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
        cur, bct = torch.from_numpy(x), torch.from_numpy(bc)
        for _ in range(args.warmup):
            cur = it(cur, bct, None)
        times = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            cur = it(cur, bct, None)
            times.append(time.perf_counter() - t0)
    print(json.dumps({"seconds_per_iteration": times, "threads": torch.get_num_threads(), "cores": os.cpu_count(),
                      "torch": torch.__version__, "checksum": float(cur.double().sum()),
                      "B": args.B, "N": args.N, "layers": args.layers, "act": args.act}))


if __name__ == "__main__":
    main()
