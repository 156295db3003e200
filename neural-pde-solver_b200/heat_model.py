"""Drop-in mirror of models/heat_model.py::HeatModel (+ the save/load of
models/base_model.py): the loop drivers of the hot path.

``train`` keeps the reference's semantics exactly -- N ~ U{1..max_iter_steps},
N-1 detached iterations then ONE differentiable iteration, MSE against gt, the
same again from gt weighted by lambda_gt (heat_model.py:31-75) -- but the N-1
detached iterations are a single fused ``iterate`` call instead of a Python loop,
and the differentiable step's backward is the hand-written kernel.  For
data-parallel training pass ``world`` (a torch.distributed process group or True
for the default group): the weight gradient (9*n_layers floats) is all-reduced
over NCCL and the loss is averaged over the GLOBAL batch.
"""
import os
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from . import heat_utils as utils
from .get_iterator import get_iterator


class HeatModel(object):
    def __init__(self, opt, world=None, full_bptt=False):
        """world: None, True (default process group) or a torch.distributed group: data-parallel training.
        full_bptt: differentiate through ALL unrolled iterations (BASELINE.json config 4) instead of the
        reference's last-step-only gradient (models/heat_model.py:48-53 detaches the first N-1 steps)."""
        self.full_bptt = full_bptt
        self.nets, self.optimizers, self.schedulers = {}, {}, []
        self.iterator, self.compare_model, self.operations_ratio, self.is_train = get_iterator(opt)
        self.nets['iterator'] = self.iterator
        self.world = world
        if self.is_train:
            self.criterion_mse = nn.MSELoss()
            if opt.optimizer == 'sgd':
                self.optimizer = optim.SGD(self.iterator.parameters(), lr=opt.lr_init)
            elif opt.optimizer == 'adam':
                self.optimizer = optim.Adam(self.iterator.parameters(), lr=opt.lr_init)
            else:
                raise NotImplementedError
            self.max_iter_steps = opt.max_iter_steps
            self.max_iter_steps_from_gt = opt.max_iter_steps_from_gt
            self.lambdas = {'gt': opt.lambda_gt}

    # ---- models/base_model.py ------------------------------------------------------
    def setup(self, is_train):
        for _, net in self.nets.items():
            net.train() if is_train else net.eval()

    def save(self, ckpt_path, epoch):
        for name, net in self.nets.items():
            torch.save(net.state_dict(), os.path.join(ckpt_path, 'net_{}_{}.pth'.format(name, epoch)))

    def load(self, ckpt_path, epoch, load_optimizer=False):
        for name, net in self.nets.items():
            path = os.path.join(ckpt_path, 'net_{}_{}.pth'.format(name, epoch))
            if not os.path.exists(path):
                print('{} does not exist, ignore.'.format(path))
                continue
            ckpt = torch.load(path, map_location=lambda storage, loc: storage)
            try:
                net.load_state_dict(ckpt)
            except Exception:
                print('net_{} and checkpoint have different parameter names'.format(name))
                new_ckpt = OrderedDict()
                for ckpt_key, module_key in zip(ckpt.keys(), net.state_dict().keys()):
                    assert ckpt_key.split('.')[-1] == module_key.split('.')[-1]
                    new_ckpt[module_key] = ckpt[ckpt_key]
                net.load_state_dict(new_ckpt)

    # ---- models/heat_model.py ------------------------------------------------------
    def _device(self):
        for p in self.iterator.parameters():
            return p.device
        return torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')

    def _to(self, *tensors):
        dev = self._device()
        return [None if t is None else t.to(dev) for t in tensors]

    def _unrolled(self, start, bc, f, n_steps):
        """n_steps-1 detached iterations (fused) + one differentiable iteration; with full_bptt every
        iteration is a differentiable step (the autograd tape keeps one field per step)."""
        y = start.detach()
        if self.full_bptt:
            if hasattr(self.iterator, 'unrolled'):
                return self.iterator.unrolled(y, bc, f, n_steps)  # one autograd node, one backward call
            for _ in range(n_steps):
                y = self.iter_step(y, bc, f)
            return y
        if n_steps > 1:
            y, _, _ = self.iterator.iterate(y, bc, f, n_steps - 1)
        return self.iter_step(y, bc, f)

    def train(self, x, gt, bc, f):
        if not self.is_train:
            return {}
        x, gt, bc, f = self._to(x, gt, bc, f)
        loss_dict = {}
        if self.lambdas['gt'] < 1:
            N = np.random.randint(1, self.max_iter_steps + 1)
            y = self._unrolled(x, bc, f, N)
            loss_x = self.criterion_mse(y, gt)
        else:
            loss_x = 0
        if self.lambdas['gt'] > 0:
            M = np.random.randint(1, self.max_iter_steps_from_gt + 1)
            y_gt = self._unrolled(gt, bc, f, M)
            loss_gt = self.criterion_mse(y_gt, gt)
        else:
            loss_gt = 0
        loss = (1 - self.lambdas['gt']) * loss_x + self.lambdas['gt'] * loss_gt
        self.optimizer.zero_grad()
        loss.backward()
        self._allreduce_grads()
        self.optimizer.step()
        # the reference reads the two loss scalars back with .item() (heat_model.py:54,66)
        if self.lambdas['gt'] < 1:
            loss_dict['loss_x'] = self._global_mean(loss_x)
        if self.lambdas['gt'] > 0:
            loss_dict['loss_gt'] = self._global_mean(loss_gt)
        return {'loss': loss_dict}

    def _group(self):
        import torch.distributed as dist
        if self.world is None or not dist.is_available() or not dist.is_initialized():
            return None, None
        return dist, (None if self.world is True else self.world)

    def _allreduce_grads(self):
        """Data parallel over the batch: every rank holds B/G problems, the loss is a mean over
        the local batch, so the global gradient is the mean of the rank gradients."""
        dist, group = self._group()
        if dist is None:
            return
        grads = [p.grad for p in self.iterator.parameters() if p.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat /= dist.get_world_size(group)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def _global_mean(self, loss):
        dist, group = self._group()
        v = loss.detach()
        if dist is not None:
            v = v.clone()
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
            v /= dist.get_world_size(group)
        return v.item()

    def iter_step(self, x, bc, f):
        return self.iterator.iter_step(x, bc, f)

    def H(self, x, bc, f):
        return self.iterator.H(x.unsqueeze(1), bc).squeeze(1)

    def get_activation(self):
        return self.iterator.act

    def change_activation(self, act):
        self.iterator.act = act

    def evaluate(self, x, gt, bc, f, n_steps):
        """heat_model.py:92-128: errors of the comparison model and of the model, (B, n_steps+1) each."""
        x, gt, bc, f = self._to(x, gt, bc, f)
        if utils.is_bc_mask(x, bc):
            print('Initializing with zero')
            x = utils.initialize(x, bc, 'zero')
        starting_error = utils.l2_error(x, gt).cpu()
        results = {}
        threshold = 0.01 if self.iterator.name().startswith('UNet') else 0.002
        if self.compare_model is not None:
            fd_errors, _ = utils.calculate_errors(x, bc, f, gt, self.compare_model.iter_step, n_steps,
                                                  starting_error, threshold)
            results['Jacobi errors'] = fd_errors
        errors, x = utils.calculate_errors(x, bc, f, gt, self.iter_step, n_steps, starting_error, threshold)
        results['model errors'] = errors
        return results, x
