"""Builds libnpde_b200.so (sm_100a) in-tree with nvcc.

The shared library is the product: hand-written CUDA kernels plus the C ABI of
include/npde.h.  There is no JIT and no fallback: if the library is missing the
package refuses to work (see _lib.py).

    python neural-pde-solver_b200/build.py            # build if sources changed
    python neural-pde-solver_b200/build.py --force
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libnpde_b200.so")
SOURCES = ["npde_grid_ops.cu", "npde_conv_ops.cu", "npde_stream.cu", "npde_wide.cu", "npde_bwd_stream.cu", "npde_unet_coarse.cu", "npde_cg.cu", "npde_small.cu", "npde_api.cu"]
HEADERS = [os.path.join(CSRC, "npde_common.cuh"), os.path.join(CSRC, "npde_internal.h"), os.path.join(CSRC, "npde_f2.cuh"),
           os.path.join(ROOT, "include", "npde.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--fmad=false"]  # no implicit contraction: every fused multiply-add is written as one


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + HEADERS):
            jobs.append([nvcc] + NVCC_FLAGS + ["-c", "-o", o, s])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))

    with ThreadPoolExecutor(max_workers=6) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", LIB] + objs)  # default: static cudart, no runtime-library lookup at load time
    return LIB


def build_variant(name, defines, source="npde_stream.cu", extra_flags=()):
    """DEVELOPMENT: a tuning build of the same library with extra -D flags for ONE translation unit
    (npde_stream.cu: only the Conv3 / clamp stream kernel is instantiated, see NPDE_DEV_ONLY_CONV3)
    -> tools/_build/libnpde_<name>.so.  Loaded through the NPDE_B200_LIB hook of _lib.py by
    tools/ab_bench.py / tools/ab_wide.py; never shipped."""
    out_dir = os.path.join(ROOT, "tools", "_build")
    os.makedirs(out_dir, exist_ok=True)
    build()
    obj = os.path.join(out_dir, "%s_%s.o" % (source[:-3], name))
    dev = ["-DNPDE_DEV_ONLY_CONV3"] if source == "npde_stream.cu" else []
    cmd = [_nvcc()] + NVCC_FLAGS + dev + ["-D" + d for d in defines] + list(extra_flags) + \
          ["-c", "-o", obj, os.path.join(CSRC, source)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), r.stderr))
    objs = [obj if s == source else os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    out = os.path.join(out_dir, "libnpde_%s.so" % name)
    r = subprocess.run([_nvcc(), "-shared", "-o", out] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed: " + r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
