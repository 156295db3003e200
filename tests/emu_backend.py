"""TEST INFRASTRUCTURE: run the package against the host-compiled SIMT emulator build of the SAME .cu sources
(tests/emu/cuda_emu.h -> tests/emu/_build/libnpde_emu.so), so that kernel index logic and host schedules are
exercised without a GPU.  The product package has no switch for this: the test-suite swaps the ctypes handle
inside neural-pde-solver_b200._lib from here.  The emulator takes HOST pointers, hence `_emulated = True`
(`_lib.host_pointers()`), which is the only thing that makes the wrappers accept CPU tensors."""
import importlib
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor


def build_emulator(force=False):
    """The product's .cu sources compiled for the host against tests/emu/cuda_emu.h (a SIMT emulator)
    -> tests/emu/_build/libnpde_emu.so."""
    build = importlib.import_module("neural-pde-solver_b200.build")
    ROOT, CSRC, SOURCES, HEADERS, _stale = build.ROOT, build.CSRC, build.SOURCES, build.HEADERS, build._stale
    emu_dir = os.path.join(ROOT, "tests", "emu")
    out = os.path.join(emu_dir, "_build", "libnpde_emu.so")
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    if force or _stale(out, srcs + HEADERS + [os.path.join(emu_dir, "cuda_emu.h")]):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        flags = ["g++", "-O2", "-std=c++17", "-DNPDE_EMU", "-ffp-contract=off", "-fPIC", "-I" + emu_dir]
        objs = [os.path.join(os.path.dirname(out), os.path.basename(s_)[:-3] + ".emu.o") for s_ in srcs]

        def cc(pair):
            src, obj = pair
            r = subprocess.run(flags + ["-x", "c++", "-c", src, "-o", obj], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("emulator build failed:\n" + r.stderr)

        with ThreadPoolExecutor(max_workers=6) as ex:  # the translation units in parallel
            list(ex.map(cc, zip(srcs, objs)))
        r = subprocess.run(["g++", "-shared", "-o", out] + objs, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("emulator link failed:\n" + r.stderr)
    return out


def use_emulator(npde=None):
    npde = npde or importlib.import_module("neural-pde-solver_b200")
    lib = npde._lib
    lib._lib = lib._bind(build_emulator())
    lib._emulated = True
    return npde


def use_cuda_library(npde=None):
    npde = npde or importlib.import_module("neural-pde-solver_b200")
    npde._lib._lib = None  # re-bound lazily to libnpde_b200.so by _lib.lib()
    npde._lib._emulated = False
    return npde
