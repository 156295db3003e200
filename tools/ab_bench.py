#!/usr/bin/env python
"""DEVELOPMENT: A/B timing of tuning builds of the Conv3 stream kernel (tools/_build/libnpde_<name>.so,
made by build.build_variant) on the bench workload (64 x 1025^2, clamp, 100 iterations).

    python tools/ab_bench.py name1 name2 ...      # on the GPU box; prints G cell-updates/s per build
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import importlib, sys, numpy as np, torch
sys.path.insert(0, %(root)r)
npde = importlib.import_module("neural-pde-solver_b200")
B, N, K = 64, 1025, 100
rng = np.random.RandomState(0)
bc = torch.from_numpy(rng.rand(B, 4).astype(np.float32)).cuda()
x = torch.from_numpy(rng.rand(B, N, N).astype(np.float32)).cuda()
x = npde.utils.set_boundary(x, bc)
it = npde.ConvIterator("clamp", 3).cuda()
with torch.no_grad():
    for l in it.layers:
        l.weight.copy_(torch.from_numpy(((rng.rand(3, 3) - 0.5) * 0.2).astype(np.float32)).view(1, 1, 3, 3))
    out = torch.empty_like(x)
    for _ in range(3):
        it.iterate(x, bc, None, K, out=out)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); it.iterate(x, bc, None, K, out=out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
print("%%.1f G cu/s  (%%.1f us / launch)  checksum %%.9g" %% (B * N * N * K / best / 1e6, best * 1e3 / K, float(out.double().sum())))
'''


def main():
    for name in sys.argv[1:]:
        lib = os.path.join(ROOT, "tools", "_build", "libnpde_%s.so" % name) if name != "default" else ""
        env = dict(os.environ)
        if lib:
            env["NPDE_B200_LIB"] = lib
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=env, capture_output=True, text=True)
        print("%-24s %s" % (name, (r.stdout.strip() or r.stderr.strip()[-400:])), flush=True)


if __name__ == "__main__":
    main()
