"""Solve-to-tolerance dataset pipeline -- the B200 mirror of generation.py.

What the reference does per run (generation.py:117-241): draw a batch of problems on the host,
warm-start with 400 iterations of a multigrid / pretrained U-Net iterator, run Jacobi until the
largest ``fd_error`` of the batch drops below 1e-5 (checked every 100 iterations, at most 8000),
and save ``[x0, solution(, f | bc_values, bc_mask)]`` as float64 scaled by ``max_temp``.

Here the two loops are device loops: the warm start is ONE ``npde_iterate`` call and the Jacobi
solve is one temporally blocked ``npde_iterate`` call per 100 iterations with a single scalar read
back (``utils.solve_to_tolerance``).  Host work (random draws, file format) follows the reference
draw for draw, so a seeded run writes the same files -- tests/test_pipeline.py compares them with
files written by the reference itself.

On-disk format (readers: data.py, runtime.get_data):
    <save_dir>/<geometry>/[poisson_]NxN/opt.npy            pickled argparse.Namespace
    .../frames/{run:04d}.npy   float64 (batch, 2|3|4, N, N)  [x0, solution] (+ f | + bc_values, bc_mask)
    .../bc.npy                 float64 (n_runs, batch, 4)    square geometry only
"""
import argparse
import os

import numpy as np
import torch

from . import heat_utils as utils
from .geometries import gaussian, get_geometry
from .iterators import JacobiIterator, MultigridIterator

ERROR_THRESHOLD = 0.00001  # generation.py:78
MAX_ITERS = 8000           # generation.py:79
WARM_START_ITERS = 400     # generation.py:162,222


def build_parser():
    """generation.py:11-29 plus --device / --init_ckpt / --seed."""
    parser = argparse.ArgumentParser()
    parser.add_argument('--save_dir', type=str, default=os.path.join(os.environ.get('HOME', '.'), 'slowbro/PDE/heat'))
    parser.add_argument('--batch_size', type=int, default=16)
    parser.add_argument('--save_every', type=int, default=1)
    parser.add_argument('--n_frames', type=int, default=1)
    parser.add_argument('--n_runs', type=int, default=1000)
    parser.add_argument('--image_size', type=int, default=17)
    parser.add_argument('--max_temp', type=int, default=100)
    parser.add_argument('--poisson', type=int, default=0)
    parser.add_argument('--geometry', type=str, default='square',
                        choices=['square', 'cylinders', 'Lshape', 'centered_cylinders', 'centered_Lshape'])
    parser.add_argument('--use_model', type=int, default=1, help='Warm-start with multigrid or a pretrained model.')
    parser.add_argument('--device', type=str, default='cuda')
    parser.add_argument('--init_ckpt', type=str, default='',
                        help='directory of a trained HeatModel (opt.npy + net_iterator_19.pth) used for the warm start')
    parser.add_argument('--seed', type=int, default=666)  # generation.py:31
    return parser


def _device(opt):
    return torch.device(getattr(opt, 'device', None) or 'cuda')


def setup_model(opt):
    """generation.py:34-70: the warm-start iterator.  Multigrid(2|4|6, 8, 8) by grid size, or the
    pretrained U-Net checkpoint if ``opt.init_ckpt`` names one (the reference hard-codes its path)."""
    ckpt = getattr(opt, 'init_ckpt', '')
    if ckpt and opt.image_size >= 64:
        from .heat_model import HeatModel
        model_opt = np.load(os.path.join(ckpt, 'opt.npy'), allow_pickle=True).item()
        model_opt.is_train = False
        model_opt.geometry = opt.geometry
        model = HeatModel(model_opt)
        model.load(ckpt, 19)
        print('Model loaded from {}'.format(ckpt))
        model.iterator.to(_device(opt))
    else:
        depth = 2 if opt.image_size < 64 else (4 if opt.image_size < 1024 else 6)
        model = MultigridIterator(depth, 8, 8)
    if opt.geometry != 'square':
        model.is_bc_mask = True
    return model


def _warm_start(model, x, bc, f, n=WARM_START_ITERS):
    fused = getattr(model, 'iterate', None) or getattr(getattr(model, 'iterator', None), 'iterate', None)
    if fused is not None:
        return fused(x, bc, f, n)[0]
    for _ in range(n):
        x = model.iter_step(x, bc, f).detach()
    return x


def get_solution(x, bc, f, verbose=True):
    """generation.py:72-99: Jacobi until the largest residual of the batch is below 1e-5.
    Returns (solution as a float32 numpy array, iterations applied)."""
    jacobi = JacobiIterator()
    res = utils.solve_to_tolerance(jacobi, x, bc, f, tol=ERROR_THRESHOLD, check_every=100, max_iters=MAX_ITERS)
    if verbose:
        for it, largest in res['history']:
            print('largest error {}'.format(largest) if it == 0 else 'Iter {}: largest error {}'.format(it, largest))
    return res['x'].cpu().numpy(), res['iterations']


def get_heat_source(image_size, batch_size, device='cuda'):
    """generation.py:101-115: negative Gaussian source, zero on the boundary ring; (B, N, N) float32."""
    heat_src = np.zeros((batch_size, image_size, image_size), np.float32)
    for i in range(batch_size):
        scale = np.random.uniform(3, 3.5) / (image_size ** 2)
        heat_src[i, 1:-1, 1:-1] = (-gaussian(image_size - 2) * scale).astype(np.float32)
    return torch.from_numpy(heat_src).to(device)


def _set_square_boundary_np(x, bc):
    """utils.set_boundary on host arrays (the reference calls it on float64 numpy, generation.py:144)."""
    x[:, 0, :] = bc[:, 0:1]
    x[:, -1, :] = bc[:, 1:2]
    x[:, :, 0] = bc[:, 2:3]
    x[:, :, -1] = bc[:, 3:4]
    return x


def generate_square(opt):
    """generation.py:117-197.  Values are scaled by max_temp at the end, like the reference."""
    dev = _device(opt)
    dir_name = ('poisson_{}x{}' if opt.poisson else '{}x{}').format(opt.image_size, opt.image_size)
    opt.save_dir = os.path.join(opt.save_dir, opt.geometry, dir_name)
    frame_dir = os.path.join(opt.save_dir, 'frames')
    os.makedirs(frame_dir, exist_ok=True)

    boundary_conditions = np.random.rand(opt.n_runs, opt.batch_size, 4)
    if opt.poisson:
        boundary_conditions *= 0.75
    np.save(os.path.join(opt.save_dir, 'opt.npy'), opt)
    model = setup_model(opt) if opt.use_model else None

    total_iters = []
    for run in range(opt.n_runs):
        x_host = _set_square_boundary_np(np.random.rand(opt.batch_size, opt.image_size, opt.image_size),
                                         boundary_conditions[run])
        original_x = x_host.copy()[:, np.newaxis, :, :]
        f = get_heat_source(opt.image_size, opt.batch_size, dev) if opt.poisson else None

        x = torch.Tensor(x_host)
        bc = torch.Tensor(boundary_conditions[run])
        # interior starts at the mean boundary temperature; the mean is taken on the host so that its
        # rounding is the reference's (generation.py:158)
        x[:, 1:-1, 1:-1] = bc.mean(dim=1).view(-1, 1, 1)
        x, bc = x.to(dev), bc.to(dev)

        if model is not None:
            x = _warm_start(model, x, bc, f)
        solution, iters = get_solution(x, bc, f)
        total_iters.append(iters)
        frames = np.concatenate([original_x, solution[:, np.newaxis].astype(np.float64)], axis=1)
        assert frames.shape[1] == opt.n_frames + 1
        if f is not None:
            frames = np.concatenate([frames, f.cpu().numpy()[:, np.newaxis].astype(np.float64)], axis=1)

        # solutions above 1 (possible with a source) are rescaled to 0.99 together with their bc
        max_value = frames[:, 1].reshape(opt.batch_size, -1).max(axis=1)
        eps = 1e-6
        over = max_value > 1 + eps
        if np.any(over):
            scaling = np.ones(opt.batch_size)
            scaling[over] = 0.99 / max_value[over]
            frames *= scaling[:, np.newaxis, np.newaxis, np.newaxis]
            boundary_conditions[run] *= scaling[:, np.newaxis]
        assert np.all(frames[:, :2, :, :] <= 1 + eps)

        np.save(os.path.join(frame_dir, '{:04d}.npy'.format(run)), frames * opt.max_temp)
        print('Run {} saved.'.format(run))

    boundary_conditions *= opt.max_temp
    np.save(os.path.join(opt.save_dir, 'bc.npy'), boundary_conditions)
    return total_iters


def generate_geometry(opt):
    """generation.py:199-241."""
    assert opt.poisson == 0
    dev = _device(opt)
    opt.save_dir = os.path.join(opt.save_dir, opt.geometry, '{}x{}'.format(opt.image_size, opt.image_size))
    frame_dir = os.path.join(opt.save_dir, 'frames')
    os.makedirs(frame_dir, exist_ok=True)
    np.save(os.path.join(opt.save_dir, 'opt.npy'), opt)
    model = setup_model(opt) if opt.use_model else None

    total_iters = []
    for run in range(opt.n_runs):
        x_host, bc_values, bc_mask = get_geometry(opt.geometry, opt.image_size, opt.batch_size, 1)
        x = torch.Tensor(x_host)
        original_x = x.numpy().copy()[:, np.newaxis, :, :]  # float32, like the reference (generation.py:215)
        bc = torch.Tensor(np.stack([bc_values, bc_mask], axis=1)).to(dev)
        x = x.to(dev)
        if model is not None:
            x = _warm_start(model, x, bc, None)
        solution, iters = get_solution(x, bc, None)
        total_iters.append(iters)
        data = np.concatenate([original_x, solution[:, np.newaxis],
                               bc_values[:, np.newaxis, :, :], bc_mask[:, np.newaxis, :, :]], axis=1)
        data[:, :3] *= opt.max_temp
        np.save(os.path.join(frame_dir, '{:04d}.npy'.format(run)), data)
        print('Run {} saved.'.format(run))
    return total_iters


def main(argv=None):
    opt = build_parser().parse_args(argv)
    assert (opt.image_size - 1) % 16 == 0, 'image_size must be 2^n + 1'
    np.random.seed(opt.seed)
    if opt.geometry == 'square':
        generate_square(opt)
    else:
        generate_geometry(opt)


if __name__ == '__main__':
    main()
