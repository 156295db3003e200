#!/usr/bin/env python
"""bench.py -- cell-updates/s of the fused Jacobi + learned-H iterator at 1025^2.

    python bench.py --gpus N --steps K --warmup W           # this repo's CUDA path
    python bench.py --impl reference ...                     # CPU baseline arm (oracle port, all host threads)

Workload (BASELINE.json configs[2]): ConvIterator (3-layer H, clamp) + Jacobi on the
square Dirichlet problem, N = 1025, batch 64 PER GPU (weak scaling: independent
problems, no data-path collective), f = None, synthetic random fields / bc /
weights.  One "step" = one pass of the reference's solve loop body
(generation.py:86-93): 100 applications of Phi_H followed by one fd_error
residual check.  1 cell-update = Phi_H applied to one grid cell (all B*N^2 cells
count), so a step is B*N^2*100 cell-updates.

value : device-resident throughput (inputs already in HBM), CUDA events.
e2e   : same step through the public API with HOST buffers: pinned x, bc copied
        host->device, iterate + residual, solution and residual copied back.
Inputs (269 MB per field) exceed the 126 MB L2, so no explicit flush is needed.
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GRID_N = 1025
BATCH_PER_GPU = 64
N_LAYERS = 3
ACT = "clamp"
INNER = 100  # iterations per step (generation.py:89 checks the residual every 100 iterations)
METRIC = "cell-updates/s (Jacobi+learned H, Conv3) at 1025^2"
UNIT = "cell-updates/s"
ALGO_BYTES_PER_CELL_UPDATE = 8.0  # read x + write y, fp32 (square, f = None)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_problem(B, N, seed):
    rng = np.random.RandomState(seed)
    bc = rng.rand(B, 4).astype(np.float32)
    x = rng.rand(B, N, N).astype(np.float32)
    x[:, 0, :] = bc[:, 0:1]
    x[:, -1, :] = bc[:, 1:2]
    x[:, :, 0] = bc[:, 2:3]
    x[:, :, -1] = bc[:, 3:4]
    w = ((rng.rand(N_LAYERS, 3, 3) - 0.5) * 0.2).astype(np.float32)
    return x, bc, w


class ClockSampler(object):
    """SM clock and throttle reasons sampled through NVML every 5 ms while the timed region runs
    (nvidia-smi's own loop mode is too slow for a region of ~100 ms)."""

    def __init__(self, index):
        self.index = index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.error = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # noqa: BLE001 -- NVML missing: report it, do not fail the bench
            self.error = "nvml unavailable: %r" % (exc,)
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        nv = self.nv
        names = [("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)]
        n = 0
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                if n % 4 == 0:  # throttle reasons: a slower query, every 4th sample
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                    for name, bit in names:
                        if mask & bit:
                            self.reasons.add(name)
            except Exception as exc:  # noqa: BLE001
                self.error = repr(exc)
                return
            n += 1
            time.sleep(0.002)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": [self.error or "not started"]}
        self._stop.set()
        self.thread.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.sm), "reasons": sorted(self.reasons)}


def cpu_oracle_throughput(x, bc, w, budget_s):
    """Oracle port (oracle/npde_oracle.c, OpenMP over the batch) on the host cores: bounded sample."""
    import oracle
    oracle.build()
    t0 = time.perf_counter()
    y, _, _ = oracle.iterate("conv", x, bc, None, w, ACT, 1)  # warm-up, also sizes the sample
    t1 = time.perf_counter() - t0
    k = int(max(1, min(50, budget_s / max(t1, 1e-3))))
    t0 = time.perf_counter()
    oracle.iterate("conv", y, bc, None, w, ACT, k)
    dt = time.perf_counter() - t0
    B, N, _ = x.shape
    return B * N * N * k / dt, k, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    oracle.build()
    cores = oracle.set_threads(os.cpu_count() or 1)  # all host cores, whatever OMP_NUM_THREADS torchrun exported
    B, N = BATCH_PER_GPU, GRID_N
    x, bc, w = make_problem(B, N, 666)
    cur = x
    for _ in range(args.warmup):
        cur, _, _ = oracle.iterate("conv", cur, bc, None, w, ACT, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cur, _, _ = oracle.iterate("conv", cur, bc, None, w, ACT, 1)
    dt = time.perf_counter() - t0
    value = B * N * N * args.steps / dt
    sample = "per step: 1 of the %d Phi_H iterations of a step, on the full %dx%d^2 batch" % (INNER, B, N)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "conv3_jacobi_square_1025_b64", "grid": N, "batch_per_gpu": B, "n_layers": N_LAYERS,
                   "activation": ACT, "iterations_per_step": 1, "note": "CPU arm: oracle port of the reference "
                   "iterator (the reference itself is Python/PyTorch and cannot travel to the GPU box)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_gpu(args):
    import torch
    import torch.distributed as dist
    npde = importlib.import_module("neural-pde-solver_b200")
    lib = npde._lib.lib()  # raises if libnpde_b200.so is missing: there is no fallback

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, N = BATCH_PER_GPU, GRID_N
    x_np, bc_np, w_np = make_problem(B, N, 666 + rank)
    it = npde.ConvIterator(ACT, N_LAYERS).to(dev)
    with torch.no_grad():
        for l, wl in zip(it.layers, w_np):
            l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
    x_host = torch.from_numpy(x_np).pin_memory()
    bc_host = torch.from_numpy(bc_np).pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()
    err_host = torch.empty(B).pin_memory()
    x_dev = x_host.to(dev)
    bc_dev = bc_host.to(dev)
    out_dev = torch.empty_like(x_dev)
    units_per_step = float(B) * N * N * INNER

    def step_resident():
        y, _, _ = it.iterate(x_dev, bc_dev, None, INNER, out=out_dev)
        return npde.utils.fd_error(y, bc_dev, None)

    # end to end: every step copies its inputs from pinned host memory and its result (solution + residual)
    # back.  Consecutive steps are software-pipelined on three streams with double-buffered device tensors --
    # H2D of step s+1 and D2H of step s-1 overlap the compute of step s -- which is how a caller feeding a stream
    # of batches through the public API (every call takes the current CUDA stream) would drive it.
    copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    x_bufs = [torch.empty_like(x_dev) for _ in range(2)]
    bc_bufs = [torch.empty_like(bc_dev) for _ in range(2)]
    out_bufs = [torch.empty_like(x_dev) for _ in range(2)]
    err_bufs = [torch.empty(B, device=dev) for _ in range(2)]

    def run_e2e(n_steps):
        compute = torch.cuda.current_stream(dev)
        h2d_done = [torch.cuda.Event() for _ in range(2)]
        comp_done = [torch.cuda.Event() for _ in range(2)]
        d2h_done = [torch.cuda.Event() for _ in range(2)]
        for s_ in range(n_steps):
            i = s_ % 2
            with torch.cuda.stream(copy_in):
                if s_ >= 2:
                    copy_in.wait_event(comp_done[i])  # step s-2 has consumed this input buffer
                x_bufs[i].copy_(x_host, non_blocking=True)
                bc_bufs[i].copy_(bc_host, non_blocking=True)
                h2d_done[i].record(copy_in)
            compute.wait_event(h2d_done[i])
            if s_ >= 2:
                compute.wait_event(d2h_done[i])  # step s-2's result has left this output buffer
            y, _, _ = it.iterate(x_bufs[i], bc_bufs[i], None, INNER, out=out_bufs[i])
            err_bufs[i].copy_(npde.utils.fd_error(y, bc_bufs[i], None))
            comp_done[i].record(compute)
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(comp_done[i])
                out_host.copy_(out_bufs[i], non_blocking=True)
                err_host.copy_(err_bufs[i], non_blocking=True)
                d2h_done[i].record(copy_out)
        torch.cuda.synchronize()  # the caller holds every solution and residual now

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ------------------------------------------------------
    with torch.no_grad():
        for _ in range(args.warmup):
            step_resident()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        launches0 = lib.npde_debug_launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for a, b in ev:
            a.record()
            y, _, _ = it.iterate(x_dev, bc_dev, None, INNER, out=out_dev)
            b.record()
            npde.utils.fd_error(y, bc_dev, None)
        e1.record()
        barrier()
        launches = lib.npde_debug_launch_count() - launches0
        total_ms = max_over_ranks(e0.elapsed_time(e1))
        kernel_ms = sum(a.elapsed_time(b) for a, b in ev) / (args.steps * INNER)  # avg launch of the fused kernel
        clocks = sampler.stop() if rank == 0 else None

        # ---- end to end through the public API with host buffers ---------------------------
        run_e2e(max(args.warmup, 2))
        barrier()
        e0.record()
        run_e2e(args.steps)
        e1.record()
        barrier()
        e2e_ms = max_over_ranks(e0.elapsed_time(e1))

    value = units_per_step * args.steps * world / (total_ms * 1e-3)
    e2e_value = units_per_step * args.steps * world / (e2e_ms * 1e-3)
    peak, peak_src = peaks()
    algo_bytes = ALGO_BYTES_PER_CELL_UPDATE * B * N * N  # per launch of the fused kernel (one Phi_H over the batch)
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_conv3_1025_b64.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get("dram_bytes_per_launch")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "conv3_jacobi_square_1025_b64", "grid": N, "batch_per_gpu": B, "global_batch": B * world,
                   "n_layers": N_LAYERS, "activation": ACT, "iterations_per_step": INNER,
                   "step": "100 x Phi_H (npde_iterate) + fd_error", "l2": "inputs (269 MB/field) exceed the 126 MB L2",
                   "parallelism": "batch-sharded x%d, no collective" % world},
        "value_per_gpu": value / world,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "pipeline": "3 streams, double-buffered: H2D(s+1) | compute(s) | D2H(s-1); every step's copies are "
                            "inside the timed region",
                "h2d_bytes_per_step": int(x_host.numel() * 4 + bc_host.numel() * 4),
                "d2h_bytes_per_step": int(out_host.numel() * 4 + err_host.numel() * 4)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_conv_wide<3,clamp,no-f>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": kernel_ms,
                     "cell_updates_per_s_kernel": B * N * N / (kernel_ms * 1e-3)},
    }
    if world == 1 and not args.no_ttr:
        # BASELINE's second metric, "time-to-1e-6 residual", on the same batch shape: plain Jacobi from the reference's
        # 'avg' start (test.py:115) until max_b fd_error < 1e-6, checked every 2000 iterations (generation.py:72-99
        # with the tolerance tightened); identical iteration count to the reference by construction (bit-identical
        # iterates, max-norm criterion).  Outside every timed region above.
        xs = npde.utils.initialize(x_dev.clone(), bc_dev, "avg")
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        res = npde.utils.solve_to_tolerance(npde.JacobiIterator(), xs, bc_dev, None, tol=1e-6, check_every=2000,
                                            max_iters=args.ttr_max_iters)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        line["time_to_residual"] = {"value": dt, "unit": "s", "tolerance": 1e-6, "iterator": "jacobi", "init": "avg",
                                    "iterations": int(res["iterations"]), "converged": bool(res["converged"]),
                                    "final_residual": float(res["history"][-1][1]),
                                    "cell_updates_per_s": B * N * N * float(res["iterations"]) / dt}
    if world == 1 and not args.no_cpu:
        import oracle
        oracle.build()
        cores = oracle.set_threads(os.cpu_count() or 1)
        v, k, dt = cpu_oracle_throughput(x_np, bc_np, w_np, args.cpu_budget)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d Phi_H iterations on the full %dx%d^2 batch (%.1f s)" % (k, B, N, dt)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ttr", action="store_true", help="skip the time-to-1e-6-residual leg (about 7 s on a B200)")
    ap.add_argument("--ttr-max-iters", type=int, default=600000)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
