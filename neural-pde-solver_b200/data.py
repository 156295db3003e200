"""Readers of the on-disk problem format written by generation.py -- mirror of
data/heat_data.py:8-149 and data/get_data_loader.py:5-26 (the data format on the
input side of the hot path).

Square geometry:   bc.npy (n_runs, batch, 4) + frames/{run:04d}.npy (batch, 2|3, N, N), all times max_temp
Other geometries:  frames/{run:04d}.npy (batch, 4, N, N) = [x0, solution, bc_values, bc_mask], first three times max_temp
Samples are dicts of float32 CPU tensors ('bc', 'x', 'final'[, 'f']); the training driver moves them to the GPU.
"""
import os
import random

import numpy as np
import torch
import torch.utils.data as data

# how a 90-degree counter-clockwise rotation of the field permutes bc = [top, bottom, left, right]
_BC_AFTER_ROT90 = {1: [3, 2, 0, 1], 2: [1, 0, 3, 2], 3: [2, 3, 1, 0]}


def _frame_path(root, i):
    return os.path.join(root, 'frames', '{:04d}.npy'.format(i))


def make_dataset(root, is_train, max_temp, data_limit, poisson):
    """data/heat_data.py:8-39: 9-1 train/test split over runs (Poisson sets are test only)."""
    bc = np.load(os.path.join(root, 'bc.npy'))
    bc /= max_temp
    total = len(bc)
    split = int(0.9 * total)
    if poisson:
        indices = np.arange(total)
    elif is_train:
        bc, indices = bc[:split], np.arange(split)
    else:
        bc, indices = bc[split:], np.arange(split, total)
    if data_limit > 0:
        bc, indices = bc[:data_limit], indices[:data_limit]

    frames_per_run = []
    for i in indices:
        frames = np.load(_frame_path(root, i))
        frames /= max_temp
        assert frames.shape[1] == (3 if poisson else 2)
        frames_per_run.append(frames)
    return bc, frames_per_run


class HeatDataset(data.Dataset):
    """data/heat_data.py:41-98."""

    def __init__(self, root, is_train, max_temp, data_limit, poisson):
        self.bc, self.data = make_dataset(root, is_train, max_temp, data_limit, poisson)
        self.n_instances, self.batch_size, _ = self.bc.shape
        self.is_train = is_train
        self.poisson = poisson

    def rotate(self, bc, x, final, f):
        k = random.randint(0, 3)
        if k > 0:
            x = np.rot90(x, k).copy()
            final = np.rot90(final, k).copy()
            if f is not None:
                f = np.rot90(f, k).copy()
            bc = bc[_BC_AFTER_ROT90[k]]
        return bc, x, final, f

    def __getitem__(self, idx):
        run, j = divmod(idx, self.batch_size)
        bc = self.bc[run, j]
        frames = self.data[run][j]
        x, final = frames[0], frames[1]
        f = frames[2] if self.poisson else None
        if self.is_train:
            bc, x, final, f = self.rotate(bc, x, final, f)
        sample = {'bc': torch.Tensor(bc), 'final': torch.Tensor(final), 'x': torch.Tensor(x)}
        if f is not None:
            sample['f'] = torch.Tensor(f)
        return sample

    def __len__(self):
        return self.n_instances * self.batch_size


def make_geometry_dataset(root, is_train, max_temp, data_limit):
    """data/heat_data.py:101-113 (no train/test split)."""
    n_instances = len(os.listdir(os.path.join(root, 'frames')))
    if 0 < data_limit < n_instances:
        n_instances = data_limit
    runs = []
    for i in range(n_instances):
        frames = np.load(_frame_path(root, i))
        frames[:, :3] /= max_temp
        runs.append(frames)
    return runs


class HeatGeometryDataset(data.Dataset):
    """data/heat_data.py:115-149."""

    def __init__(self, root, is_train, max_temp, data_limit):
        self.data = make_geometry_dataset(root, is_train, max_temp, data_limit)
        self.n_instances = len(self.data)
        self.batch_size = self.data[0].shape[0]
        self.is_train = is_train

    def rotate(self, frames):
        k = random.randint(0, 3)
        if k > 0:
            frames = np.rot90(frames, k, axes=(1, 2)).copy()
        return frames

    def __getitem__(self, idx):
        run, j = divmod(idx, self.batch_size)
        frames = self.data[run][j]
        if self.is_train:
            frames = self.rotate(frames)
        return {'x': torch.Tensor(frames[0]), 'final': torch.Tensor(frames[1]), 'bc': torch.Tensor(frames[2:])}

    def __len__(self):
        return self.n_instances * self.batch_size


def get_dataset(opt):
    """data/get_data_loader.py:5-10."""
    if opt.geometry == 'square':
        return HeatDataset(opt.dset_path, opt.is_train, opt.max_temp, opt.data_limit, opt.poisson)
    return HeatGeometryDataset(opt.dset_path, opt.is_train, opt.max_temp, opt.data_limit)


def get_data_loader(opt):
    """data/get_data_loader.py:12-16."""
    return data.DataLoader(get_dataset(opt), batch_size=opt.batch_size, shuffle=opt.is_train,
                           num_workers=opt.n_workers, pin_memory=torch.cuda.is_available())


def get_random_data_loader(opt):
    """data/get_data_loader.py:18-26: random batches for evaluation during training."""
    dset = get_dataset(opt)
    return data.DataLoader(dset, batch_size=opt.batch_size, sampler=data.RandomSampler(dset))
