"""The published-number protocol of the reference -- runtime.py:10-118 -- on the device
(SURVEY.md 8f.2): for each of `num` stored problems find the number of iterations the
model needs to reach 1 % relative RMS error against the stored solution, then time that
many applications of the iterator; the sum over problems is what
visualizations/plot_curves.py:29-46 plots ("time for 100 solves").

Differences from the reference are in the loops only:
  * the step count comes from ``utils.calculate_errors`` (fused per-step error rows);
  * the timed region is ONE ``iterate(k=steps)`` call bracketed by CUDA events (the reference
    brackets a Python loop of ``iter_step`` calls with time.time(), runtime.py:76-80);
  * ``runtime_batched`` additionally solves all problems as one batch (what a B200 is for) and
    reports the time until the LAST problem is below the threshold.
"""
import os
import time

import numpy as np
import torch

from . import heat_utils as utils

THRESHOLD = 0.01  # runtime.py:66


def get_data(geometry, num, dset_path, max_temp):
    """runtime.py:10-39: the first `num` problems of a generated set as CPU tensors."""
    n = num // 16 + 1
    runs = []
    for i in range(n):
        frames = np.load(os.path.join(dset_path, 'frames', '{:04d}.npy'.format(i)))
        if geometry == 'square':
            frames = frames / max_temp
        else:
            frames[:, :3] /= max_temp
        runs.append(frames)
    frames = torch.Tensor(np.concatenate(runs, axis=0)[:num])
    if geometry == 'square':
        bc = np.load(os.path.join(dset_path, 'bc.npy'))[:n] / max_temp
        bc = torch.Tensor(bc.reshape((-1, 4))[:num])
    else:
        bc = frames[:, 2:]
    return {'bc': bc, 'x': frames[:, 0], 'final': frames[:, 1]}


def _timed(fn, device):
    """Seconds taken by fn() on `device` (CUDA events on the current stream; wall clock for host tensors)."""
    if device.type == 'cuda':
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        start.record()
        fn()
        stop.record()
        stop.synchronize()
        return start.elapsed_time(stop) * 1e-3
    t0 = time.time()
    fn()
    return time.time() - t0


def _iterate_fn(model):
    it = getattr(model, 'iterator', model)
    return it.iterate


def steps_to_threshold(opt, model, x, bc, gt, threshold=THRESHOLD):
    """runtime.py:56-74 for a batch: per problem the first iteration with relative error < threshold
    (-1 = never within opt.n_evaluation_steps), and the initialised start field."""
    y = utils.initialize(x.clone(), bc, 'zero')
    starting_error = utils.l2_error(y, gt).cpu()
    x = utils.initialize(x, bc, opt.initialization)
    B = x.shape[0]
    first = np.full(B, -1, dtype=np.int64)
    # calculate_errors stops once ALL problems are below its threshold and pads with zeros: exact for B = 1
    # (the reference's case); for a batch the per-problem first crossing needs every row, so disable the early exit
    stop_at = threshold if B == 1 else float('-inf')
    errors, _ = utils.calculate_errors(x, bc, None, gt, model.iter_step, opt.n_evaluation_steps,
                                       starting_error, stop_at, verbose=False)
    below = errors.numpy() < threshold
    has = below.any(axis=1)
    first[has] = below.argmax(axis=1)[has]
    return first, x


def runtime(opt, model, data, device=None, verbose=True):
    """runtime.py:41-85: problems one at a time (B = 1).  Returns the list of times (s) of the solved ones."""
    device = torch.device(device or 'cuda')
    iterate = _iterate_fn(model)
    times, steps_list = [], []
    for i in range(data['bc'].size(0)):
        bc = data['bc'][i].unsqueeze(0).to(device)
        gt = data['final'][i].unsqueeze(0).to(device)
        x = data['x'][i].unsqueeze(0).to(device)
        first, x = steps_to_threshold(opt, model, x, bc, gt)
        if first[0] < 0:
            if verbose:
                print('Skip')
            continue
        steps = int(first[0])
        t = _timed(lambda: iterate(x, bc, None, steps) if steps > 0 else None, device)
        if verbose:
            print('Steps:', steps)
            print('Time: {}'.format(t))
        times.append(t)
        steps_list.append(steps)
    return times, steps_list


def runtime_batched(opt, model, data, device=None):
    """All problems as ONE batch: time of max_b steps_b fused iterations (every problem is then below the
    threshold).  Returns (seconds, per-problem steps)."""
    device = torch.device(device or 'cuda')
    bc, gt, x = (data[k].to(device) for k in ('bc', 'final', 'x'))
    first, x = steps_to_threshold(opt, model, x, bc, gt)
    solved = first >= 0
    if not solved.any():
        return 0.0, first
    steps = int(first[solved].max())
    iterate = _iterate_fn(model)
    keep = torch.from_numpy(np.nonzero(solved)[0]).to(device)
    xs, bcs = x.index_select(0, keep).contiguous(), bc.index_select(0, keep).contiguous()
    if steps > 0:
        iterate(xs, bcs, None, min(steps, 2))  # warm-up launch (workspace allocation)
    t = _timed(lambda: iterate(xs, bcs, None, steps) if steps > 0 else None, device)
    return t, first


def report(times, num):
    """runtime.py:117-118."""
    print('{} examples, {:.3f} sec'.format(len(times), sum(times)))
    print('Adjusted time: {:.3f} sec'.format(np.mean(times) * num))
    return float(np.mean(times) * num) if len(times) else float('nan')
