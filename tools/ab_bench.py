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
import os
B, N, K, NL = int(os.environ.get("AB_B", 64)), int(os.environ.get("AB_N", 1025)), 100, int(os.environ.get("AB_L", 3))
rng = np.random.RandomState(0)
bc = torch.from_numpy(rng.rand(B, 4).astype(np.float32)).cuda()
x = torch.from_numpy(rng.rand(B, N, N).astype(np.float32)).cuda()
x = npde.utils.set_boundary(x, bc)
it = npde.ConvIterator("clamp", NL).cuda()
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
    jb = 1e9
    jit = npde.JacobiIterator()
    for d in ((1, 8) if os.environ.get("AB_JACOBI") else ()):
        jit.iterate(x, bc, None, 64, out=out, temporal_depth=d)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); jit.iterate(x, bc, None, 64, out=out, temporal_depth=d); e1.record(); torch.cuda.synchronize()
        print("   jacobi depth %%d: %%.1f G cu/s (%%.1f us per launch)" %% (d, B * N * N * 64 / e0.elapsed_time(e1) / 1e6, e0.elapsed_time(e1) * 1e3 / (64 / d)))
print("B=%%d N=%%d L=%%d: %%.1f G cu/s  (%%.1f us / launch)  checksum %%.9g" %% (B, N, NL, B * N * N * K / best / 1e6, best * 1e3 / K, float(out.double().sum())))
'''


def main():
    for name in sys.argv[1:]:
        lib = os.path.join(ROOT, "tools", "_build", "libnpde_%s.so" % name) if name != "default" else ""
        env = dict(os.environ)
        if lib:
            env["NPDE_B200_LIB"] = lib
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=env, capture_output=True, text=True)
        out = r.stdout.strip().splitlines()
        if not out:
            out = [r.stderr.strip()[-400:]]
        if os.environ.get("AB_VERBOSE"):
            for line in out[:-1]:
                print("    " + line.strip())
        print("%-24s %s" % (name, out[-1]), flush=True)


if __name__ == "__main__":
    main()
