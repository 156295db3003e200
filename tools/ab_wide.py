#!/usr/bin/env python
"""DEVELOPMENT: A/B timing of the Phi_H kernels on the bench workload (Conv3, clamp, 100 iterations per call).

    python tools/ab_wide.py spec1 spec2 ...        # on the GPU box
    spec = libname[:ENV=VALUE[,ENV=VALUE...]]      libname "default" = the shipped library, else
                                                   tools/_build/libnpde_<libname>.so (build.build_variant)
    AB_B / AB_N / AB_L / AB_K select the workload (default 64 x 1025^2, 3 layers, 100 iterations).
Prints G cell-updates/s, microseconds per Phi_H application and a checksum per spec (best of 5 calls).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import importlib, sys, os, numpy as np, torch
sys.path.insert(0, %(root)r)
npde = importlib.import_module("neural-pde-solver_b200")
B, N, K, NL = (int(os.environ.get(k, d)) for k, d in (("AB_B", 64), ("AB_N", 1025), ("AB_K", 100), ("AB_L", 3)))
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
    ts = []
    for _ in range(int(os.environ.get("AB_REPS", 5))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); it.iterate(x, bc, None, K, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
best, last = min(ts), ts[-1]
print("B=%%d N=%%d L=%%d: %%.1f G cu/s  (%%.1f us / step best, %%.1f last)  checksum %%.9g" %% (
    B, N, NL, B * N * N * K / best / 1e6, best * 1e3 / K, last * 1e3 / K, float(out.double().sum())))
'''


def main():
    for spec in sys.argv[1:]:
        name, _, envs = spec.partition(":")
        env = dict(os.environ)
        if name != "default":
            env["NPDE_B200_LIB"] = os.path.join(ROOT, "tools", "_build", "libnpde_%s.so" % name)
        for kv in filter(None, envs.split(",")):
            k, _, v = kv.partition("=")
            env[k] = v
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=env, capture_output=True, text=True)
        out = r.stdout.strip().splitlines() or [r.stderr.strip()[-400:]]
        print("%-36s %s" % (spec, out[-1]), flush=True)


if __name__ == "__main__":
    main()
