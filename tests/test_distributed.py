"""Multi-process (world_size 2, gloo, CPU) tests of the N > 1 path.

The hot path shards by independent problems: every rank iterates its own slice of the batch with
NO data-path collective, and only the training step exchanges the 9*n_layers weight-gradient
floats (all-reduce).  These tests run the same host logic the GPU ranks run (bench.py --gpus N,
HeatModel(opt, world=True)) on the CPU emulator build of the kernels:

  * sharded iterate == the matching slice of the single-process result, bit for bit;
  * two data-parallel ranks training on half batches == one process training on the full batch.
"""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(rank, world, port):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if os.path.join(ROOT, "tests") not in sys.path:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emu_backend
    return emu_backend.use_emulator()


def _problem(B, N, seed):
    rng = np.random.RandomState(seed)
    bc = rng.rand(B, 4).astype(np.float32)
    x = rng.rand(B, N, N).astype(np.float32)
    x[:, 0, :] = bc[:, 0:1]
    x[:, -1, :] = bc[:, 1:2]
    x[:, :, 0] = bc[:, 2:3]
    x[:, :, -1] = bc[:, 3:4]
    gt = x * 0.5 + 0.25
    w = ((rng.rand(3, 3, 3) - 0.5) * 0.3).astype(np.float32)
    return x, gt, bc, w


def _make_opt():
    import argparse
    return argparse.Namespace(iterator="conv", activation="clamp", conv_n_layers=3, mg_n_layers=3, mg_pre_smoothing=2,
                              mg_post_smoothing=2, cg_n_iters=4, geometry="square", is_train=True, optimizer="sgd",
                              lr_init=1e-2, max_iter_steps=4, max_iter_steps_from_gt=3, lambda_gt=0.25)


def _load(model, w):
    with torch.no_grad():
        for l, wl in zip(model.iterator.layers, w):
            l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))


def _worker_shard(rank, world, port, out_dir):
    npde = _setup(rank, world, port)
    B, N = 6, 33
    x, gt, bc, w = _problem(B, N, 7)
    it = npde.ConvIterator("clamp", 3)
    with torch.no_grad():
        for l, wl in zip(it.layers, w):
            l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
        per = B // world
        sl = slice(rank * per, (rank + 1) * per)
        y, _, efd = it.iterate(torch.from_numpy(x[sl]), torch.from_numpy(bc[sl]), None, 7, want_fd=True)
        # the only cross-rank value of a solve: "has every problem converged" (generation.py:90-93)
        flag = torch.tensor([float(efd[-1].max())])
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    np.savez(os.path.join(out_dir, "shard%d.npz" % rank), y=y.numpy(), fd=efd.numpy(), flag=flag.numpy())
    dist.destroy_process_group()


def _worker_train(rank, world, port, out_dir):
    npde = _setup(rank, world, port)
    B, N = 4, 17
    x, gt, bc, w = _problem(B, N, 11)
    model = npde.HeatModel(_make_opt(), world=True)
    _load(model, w)
    per = B // world
    sl = slice(rank * per, (rank + 1) * per)
    np.random.seed(99)  # the unroll lengths N, M must be drawn identically on all ranks
    losses = []
    for _ in range(3):
        out = model.train(torch.from_numpy(x[sl]), torch.from_numpy(gt[sl]), torch.from_numpy(bc[sl]), None)
        losses.append([out["loss"]["loss_x"], out["loss"]["loss_gt"]])
    wts = np.stack([l.weight.detach().numpy()[0, 0] for l in model.iterator.layers])
    np.savez(os.path.join(out_dir, "train%d.npz" % rank), w=wts, losses=np.array(losses))
    dist.destroy_process_group()


def _single_process(npde_backend):
    npde = npde_backend
    B, N = 6, 33
    x, gt, bc, w = _problem(B, N, 7)
    it = npde.ConvIterator("clamp", 3)
    with torch.no_grad():
        for l, wl in zip(it.layers, w):
            l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
        y, _, efd = it.iterate(torch.from_numpy(x), torch.from_numpy(bc), None, 7, want_fd=True)
    return y.numpy(), efd.numpy()


@pytest.fixture()
def emu_npde():
    import emu_backend
    npde = emu_backend.use_emulator()
    yield npde
    emu_backend.use_cuda_library(npde)


def test_batch_sharding_world2(tmp_path, emu_npde):
    mp.spawn(_worker_shard, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    y_full, fd_full = _single_process(emu_npde)
    parts = [np.load(os.path.join(str(tmp_path), "shard%d.npz" % r)) for r in range(2)]
    y = np.concatenate([p["y"] for p in parts])
    fd = np.concatenate([p["fd"] for p in parts], axis=1)
    assert np.array_equal(y, y_full), "sharded iterate differs from the single-process result"
    assert np.array_equal(fd, fd_full)
    # both ranks hold the same global convergence flag = max over all problems
    assert parts[0]["flag"][0] == parts[1]["flag"][0] == np.float32(fd_full[-1].max())


def test_data_parallel_training_world2(tmp_path, emu_npde):
    mp.spawn(_worker_train, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "train%d.npz" % r)) for r in range(2)]
    # replicas stay identical
    assert np.array_equal(parts[0]["w"], parts[1]["w"])
    assert np.array_equal(parts[0]["losses"], parts[1]["losses"])
    # and equal the single-process run on the full batch (mean of two equal-size shard means == global
    # mean; gradient sums are re-associated across ranks: 1e-6 absolute on weights of magnitude 0.1)
    B, N = 4, 17
    x, gt, bc, w = _problem(B, N, 11)
    model = emu_npde.HeatModel(_make_opt())
    _load(model, w)
    np.random.seed(99)
    losses = []
    for _ in range(3):
        out = model.train(torch.from_numpy(x), torch.from_numpy(gt), torch.from_numpy(bc), None)
        losses.append([out["loss"]["loss_x"], out["loss"]["loss_gt"]])
    wts = np.stack([l.weight.detach().numpy()[0, 0] for l in model.iterator.layers])
    assert np.allclose(parts[0]["w"], wts, rtol=0, atol=1e-6), np.abs(parts[0]["w"] - wts).max()
    assert np.allclose(parts[0]["losses"], np.array(losses), rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("full_bptt", ["0", "1"])
def test_data_parallel_training_nccl_2gpu(full_bptt):
    """HeatModel(opt, world=True) on two B200s over NCCL (tools/dp_train_check.py under torchrun): the replicas stay
    bit-identical and match a single-process run on the whole batch to 1e-6, last-step gradient and backprop through
    all unrolled steps.  Needs two GPUs (two ranks must never share one GPU: B200_PROFILING.md)."""
    import subprocess
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, NPDE_FULL_BPTT=full_bptt)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "dp_train_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "replicas identical: True" in r.stdout, r.stdout[-2000:]
