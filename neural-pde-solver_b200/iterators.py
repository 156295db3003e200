"""Drop-in iterators: same classes, constructor arguments, attributes and
state-dict keys as models/iterators/*.py of the reference, with ``forward``
running as fused CUDA kernels through the C ABI instead of an F.conv2d chain.

Extension over the reference contract: every iterator also has
``iterate(x, bc, f, k, gt=None, want_l2=False, want_fd=False)`` -- k applications
in one library call -- which the loop drivers (heat_utils.calculate_errors,
HeatModel.train/evaluate, runtime) use instead of Python loops.
"""
import numpy as np
import torch
import torch.nn as nn

from . import functional as F_
from . import heat_utils as utils
from ._lib import IT_CONV, IT_JACOBI, IT_MULTIGRID, IT_UNET


class Iterator(nn.Module):
    """models/iterators/iterator.py:6-44."""

    def __init__(self, act=None):
        super(Iterator, self).__init__()
        self.act = act
        self.n_operations = None
        self.is_bc_mask = False

    def activation(self, x):
        if self.act == 'sigmoid':
            x = torch.sigmoid(x)
        elif self.act == 'clamp':
            x = x.clamp(0, 1)
        elif self.act == 'none':
            pass
        else:
            raise NotImplementedError
        return x

    def _check_act(self):
        if self.act not in ('sigmoid', 'clamp', 'none'):
            raise NotImplementedError  # iterator.py:25-26

    def iter_step(self, x, bc, f):
        return self.forward(x, bc, f)

    def forward(self, x, bc, f):
        raise NotImplementedError

    def iterate(self, x, bc, f, k, gt=None, want_l2=False, want_fd=False, out=None):
        raise NotImplementedError

    def name(self):
        raise NotImplementedError


class JacobiIterator(Iterator):
    """models/iterators/jacobi_iterator.py:7-20."""

    def __init__(self):
        super(JacobiIterator, self).__init__()
        self.n_operations = 1

    def forward(self, x, bc, f):
        return utils.fd_step(x, bc, f)

    def iterate(self, x, bc, f, k, gt=None, want_l2=False, want_fd=False, out=None, temporal_depth=0):
        return F_.iterate(IT_JACOBI, x, bc, f, k, gt=gt, want_l2=want_l2, want_fd=want_fd,
                          temporal_depth=temporal_depth, out=out)

    def solve(self, x, bc, f, tol, check_every, max_iters):
        """generation.py:72-99 as one device-side loop (npde_solve): see functional.solve."""
        return F_.solve(IT_JACOBI, x, bc, f, tol, check_every, max_iters)

    def name(self):
        return 'Jacobi'


class MultigridIterator(Iterator):
    """models/iterators/jacobi_iterator.py:23-78: one V-cycle per forward."""

    def __init__(self, n_layers, pre_smoothing, post_smoothing):
        super(MultigridIterator, self).__init__()
        self.n_layers = n_layers
        self.pre_smoothing = pre_smoothing
        self.post_smoothing = post_smoothing
        self.n_operations = (pre_smoothing + post_smoothing + 2) * 4 / 3

    def forward(self, x, bc, f):
        return F_.mg_step(x, bc, f, self.n_layers, self.pre_smoothing, self.post_smoothing)

    def iterate(self, x, bc, f, k, gt=None, want_l2=False, want_fd=False, out=None):
        return F_.iterate(IT_MULTIGRID, x, bc, f, k, n_layers=self.n_layers, pre=self.pre_smoothing,
                          post=self.post_smoothing, gt=gt, want_l2=want_l2, want_fd=want_fd, out=out)

    def solve(self, x, bc, f, tol, check_every, max_iters):
        return F_.solve(IT_MULTIGRID, x, bc, f, tol, check_every, max_iters, n_layers=self.n_layers,
                        pre=self.pre_smoothing, post=self.post_smoothing)

    def name(self):
        return 'Multigrid'


def _iterate_stepwise(it, x, bc, f, k, gt, want_l2, want_fd, out):
    """iterate() contract for iterators without a fused k-step kernel (comparison solvers outside the hot
    path): k forward calls; the error rows still come from the device reductions, no host sync."""
    x = x.detach()
    rows_fd = [utils.fd_error(x, bc, f)] if want_fd else None
    rows_l2 = [] if (want_l2 and gt is not None) else None
    for _ in range(k):
        x = it.forward(x, bc, f)
        if rows_fd is not None:
            rows_fd.append(utils.fd_error(x, bc, f))
        if rows_l2 is not None:
            rows_l2.append(utils.l2_error(x, gt))
    if out is not None:
        out.copy_(x)
        x = out
    return (x, torch.stack(rows_l2) if rows_l2 else None, torch.stack(rows_fd) if rows_fd is not None else None)


class MultigridResidualIterator(Iterator):
    """models/iterators/jacobi_iterator.py:81-143: V-cycle on the residual equation (the reference notes it
    "doesn't work well" near boundaries; kept for API completeness).  Composition of the grid operators of
    this package; the two field additions are element-wise torch ops (exact IEEE, same as the reference)."""

    def __init__(self, n_layers, pre_smoothing, post_smoothing):
        super(MultigridResidualIterator, self).__init__()
        self.n_layers = n_layers
        self.pre_smoothing = pre_smoothing
        self.post_smoothing = post_smoothing
        self.n_operations = (pre_smoothing + post_smoothing + 2) * 4 / 3

    def multigrid_step(self, x, bc, f, step):
        if step == 0:
            return None
        x = utils.set_boundary(x.clone(), bc)
        if self.pre_smoothing > 0:
            x = JacobiIterator().iterate(x, bc, f, self.pre_smoothing)[0]
        r = utils.fd_step(x, bc, f) - x
        zeros_bc = torch.zeros(x.shape[0], 4, device=x.device)
        r_sub = utils.restriction(r, zeros_bc)
        ek_sub = self.multigrid_step(r_sub, zeros_bc, -r_sub, step - 1)
        if ek_sub is not None:
            x = x + utils.interpolation(ek_sub, zeros_bc)
        x = utils.set_boundary(x, bc)
        if self.post_smoothing > 0:
            x = JacobiIterator().iterate(x, bc, f, self.post_smoothing)[0]
        return x

    def forward(self, x, bc, f):
        return self.multigrid_step(x.detach(), bc, f, self.n_layers)

    def iterate(self, x, bc, f, k, gt=None, want_l2=False, want_fd=False, out=None):
        return _iterate_stepwise(self, x, bc, f, k, gt, want_l2, want_fd, out)

    def name(self):
        return 'Multigrid residual'


class ConjugateGradient(Iterator):
    """models/iterators/conjugate_gradient.py:8-49: n_iters CG iterations per forward (one npde_cg_step call;
    alpha / beta stay on the device)."""

    def __init__(self, n_iters):
        super(ConjugateGradient, self).__init__()
        self.n_iters = n_iters
        self.n_operations = n_iters

    def forward(self, x, bc, f):
        if f is not None:
            # the reference adds f (B,N,N) to a (B,1,N-2,N-2) residual, which does not broadcast (:24-25)
            raise RuntimeError("ConjugateGradient supports f=None only (as in the reference)")
        return F_.cg_step(x, self.n_iters)

    def iterate(self, x, bc, f, k, gt=None, want_l2=False, want_fd=False, out=None):
        return _iterate_stepwise(self, x, bc, f, k, gt, want_l2, want_fd, out)

    def name(self):
        return 'Conjugate Gradient'


class _WeightCache(object):
    """Device (n,3,3) stack + host numpy copy of a list of (1,1,3,3) conv weights.  The host
    copy lets the fused kernel take the weights as launch parameters; it is refreshed only
    when a parameter's version counter moves (one tiny D2H per optimizer step)."""

    def __init__(self):
        self.key = None
        self.dev = None
        self.host = None

    def get(self, params, device=None):
        """Writes through `p.data` (old-style torch code) do not move the version counter: call `invalidate()`
        (Iterator.invalidate_weight_cache) after such a write.  `device`: where an EMPTY stack lives (Conv0)."""
        key = tuple((p.data_ptr(), p._version, p.device) for p in params) + (str(device),)
        if key != self.key:
            with torch.no_grad():
                if len(params):
                    self.dev = torch.stack([p.detach().reshape(3, 3) for p in params]).contiguous()
                else:
                    self.dev = torch.empty((0, 3, 3), device=device if device is not None else 'cpu')
                self.host = self.dev.cpu().numpy().copy()
            self.key = key
        return self.dev, self.host

    def invalidate(self):
        self.key = None


class ConvIterator(Iterator):
    """models/iterators/conv_iterator.py:8-54.  State-dict keys: layers.{i}.weight (1,1,3,3)."""

    def __init__(self, act, n_layers=1):
        super(ConvIterator, self).__init__(act)
        layers = []
        for i in range(n_layers):
            layers += [nn.Conv2d(1, 1, 3, stride=1, padding=1, bias=False)]
        self.layers = nn.ModuleList(layers)
        self.n_layers = n_layers
        self.n_operations = 1 + n_layers
        self._cache = _WeightCache()

    def _params(self):
        return [l.weight for l in self.layers]

    def _consistent(self, x, bc):
        if utils.is_bc_mask(x, bc) != bool(self.is_bc_mask):
            raise NotImplementedError(
                "is_bc_mask=%r does not match the bc tensor shape %s" % (self.is_bc_mask, tuple(bc.shape)))

    def H(self, r, bc):
        """r: (B,1,N-2,N-2) -> H(r), same shape (conv_iterator.py:19-31)."""
        w, _ = self._cache.get(self._params(), r.device)
        return F_.conv_H(r.squeeze(1), bc, w, self.is_bc_mask).unsqueeze(1)

    def forward(self, x, bc, f):
        self._check_act()
        self._consistent(x, bc)
        params = self._params()
        w_dev, w_host = self._cache.get(params, x.device)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
            w = torch.stack([p.reshape(3, 3) for p in params]) if params else w_dev
            return F_.ConvStepFn.apply(x, bc, f, w, self.act, w_host)
        return F_.conv_step(x, bc, f, w_dev, self.act, w_host)

    def unrolled(self, x, bc, f, k):
        """k chained forward calls as one differentiable node (backprop through all k iterations)."""
        self._check_act()
        self._consistent(x, bc)
        params = self._params()
        if k < 1 or not params:
            for _ in range(k):
                x = self.forward(x, bc, f)
            return x
        _, w_host = self._cache.get(params)
        w = torch.stack([p.reshape(3, 3) for p in params])
        return F_.ConvUnrollFn.apply(x, bc, f, w, self.act, w_host, k)

    def iterate(self, x, bc, f, k, gt=None, want_l2=False, want_fd=False, out=None, temporal_depth=0):
        """temporal_depth: 0 = library default (grids up to 65^2 run the whole loop in one on-chip kernel);
        1 forces the streaming kernel (one launch per step)."""
        self._check_act()
        self._consistent(x, bc)
        w_dev, w_host = self._cache.get(self._params(), x.device)
        return F_.iterate(IT_CONV, x, bc, f, k, w=w_dev, w_host=w_host, n_layers=self.n_layers, act=self.act,
                          gt=gt, want_l2=want_l2, want_fd=want_fd, out=out, temporal_depth=temporal_depth)

    def solve(self, x, bc, f, tol, check_every, max_iters):
        self._check_act()
        self._consistent(x, bc)
        w_dev, w_host = self._cache.get(self._params(), x.device)
        return F_.solve(IT_CONV, x, bc, f, tol, check_every, max_iters, w=w_dev, w_host=w_host,
                        n_layers=self.n_layers, act=self.act)

    # -- persistent packed layout (square domain, 1..4 layers): pack once, iterate / check many times, unpack once --
    def pack(self, x):
        """(B,N,N) -> functional.PackedField in the layout of the fused kernel (npde_pack)."""
        return F_.pack(x)

    def unpack(self, p, out=None):
        return F_.unpack(p, out=out)

    def iterate_packed(self, a, b, bc, f_packed, k):
        """k steps x <- Phi_H(x) from PackedField a, with PackedField b as the second buffer; returns the one holding
        the result (heat_utils.py:163-181 / generation.py:86-93 without a layout pass per call)."""
        self._check_act()
        if not 1 <= self.n_layers <= 4 or tuple(bc.shape) != (a.B, 4):
            raise NotImplementedError("packed fields: square domain, 1..4 layers")
        _, w_host = self._cache.get(self._params(), a.device)
        return F_.conv_iterate_packed(a, b, bc, f_packed, w_host, self.n_layers, self.act, k)

    def fd_error_packed(self, p, f_packed=None):
        """utils.fd_error(x, bc, f) ('max' over the interior) of a packed field, bit-identical to the unpacked one."""
        return F_.fd_error_packed(p, f_packed)

    def invalidate_weight_cache(self):
        """Call after writing the weights through `.data` (which does not move torch's version counter)."""
        self._cache.invalidate()

    def name(self):
        return 'Conv{}'.format(self.n_layers)


class UNetIterator(Iterator):
    """models/iterators/unet_iterator.py:8-108.  State-dict keys:
    first_layers.{i}.weight, pooling_layers.{i}.weight, second_layers.{i}.weight."""

    def __init__(self, act, n_layers, pre_smoothing, post_smoothing):
        super(UNetIterator, self).__init__(act)
        self.n_layers = n_layers
        self.pre_smoothing = pre_smoothing
        self.post_smoothing = post_smoothing
        self.n_operations = (pre_smoothing + post_smoothing + 2) * 4 / 3
        first_layers = []
        for i in range(n_layers):
            for j in range(pre_smoothing):
                first_layers.append(nn.Conv2d(1, 1, 3, stride=1, padding=1, bias=False))
        self.first_layers = nn.ModuleList(first_layers)
        pooling_layers = []
        for i in range(n_layers - 1):
            pooling_layers.append(nn.Conv2d(1, 1, 3, stride=2, padding=0, bias=False))
        self.pooling_layers = nn.ModuleList(pooling_layers)
        second_layers = []
        for i in range(n_layers):
            for j in range(post_smoothing):
                second_layers.append(nn.Conv2d(1, 1, 3, stride=1, padding=1, bias=False))
        self.second_layers = nn.ModuleList(second_layers)
        self._cache = _WeightCache()

    def _params(self):
        return ([l.weight for l in self.first_layers] + [l.weight for l in self.pooling_layers] +
                [l.weight for l in self.second_layers])

    def _split(self, w):
        nf = self.n_layers * self.pre_smoothing
        npool = self.n_layers - 1
        return w[:nf], w[nf:nf + npool], w[nf + npool:]

    def H(self, r, bc):
        w, _ = self._cache.get(self._params())
        wf, wp, ws = self._split(w)
        return F_.unet_H(r.squeeze(1), bc, wf, wp, ws, self.n_layers, self.pre_smoothing, self.post_smoothing,
                         self.is_bc_mask).unsqueeze(1)

    def forward(self, x, bc, f):
        self._check_act()
        if utils.is_bc_mask(x, bc) != bool(self.is_bc_mask):
            raise NotImplementedError("is_bc_mask does not match the bc tensor shape")
        params = self._params()
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params)):
            w = torch.stack([p.reshape(3, 3) for p in params])
            return F_.UNetStepFn.apply(x, bc, f, w, self.n_layers * self.pre_smoothing, self.n_layers - 1,
                                       self.n_layers, self.pre_smoothing, self.post_smoothing, self.act)
        w, w_host = self._cache.get(params)
        wf, wp, ws = self._split(w)
        return F_.unet_step(x, bc, f, wf, wp, ws, self.n_layers, self.pre_smoothing, self.post_smoothing, self.act,
                            w_host=w_host)

    def iterate(self, x, bc, f, k, gt=None, want_l2=False, want_fd=False, out=None):
        self._check_act()
        w, w_host = self._cache.get(self._params())
        return F_.iterate(IT_UNET, x, bc, f, k, w=w, w_host=w_host, n_layers=self.n_layers, pre=self.pre_smoothing,
                          post=self.post_smoothing, act=self.act, gt=gt, want_l2=want_l2, want_fd=want_fd, out=out)

    def solve(self, x, bc, f, tol, check_every, max_iters):
        self._check_act()
        w, w_host = self._cache.get(self._params())
        return F_.solve(IT_UNET, x, bc, f, tol, check_every, max_iters, w=w, w_host=w_host, n_layers=self.n_layers,
                        pre=self.pre_smoothing, post=self.post_smoothing, act=self.act)

    def invalidate_weight_cache(self):
        self._cache.invalidate()

    def name(self):
        return 'UNet{}'.format(self.n_layers)
