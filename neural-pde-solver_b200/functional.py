"""Tensor-level wrappers over the C ABI (include/npde.h).

Every function takes torch tensors that live on a CUDA device, hands raw device
pointers plus the current CUDA stream to libnpde_b200.so, and returns fresh
output tensors.  Nothing here computes: torch is used for memory, streams and
autograd bookkeeping only.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import ACT, AGG, BC_MASK, BC_SQUARE, IterateDesc, check

_workspaces = {}


class _Ptr(ctypes.c_void_p):
    """A raw device pointer that remembers which device it lives on (for the same-device check below)."""
    device = None


class _Guarded(object):
    """The C library launches on the process's CURRENT device (kernels, capture streams, occupancy queries),
    whereas the reference's torch ops run wherever their tensors live.  Every C-ABI call therefore goes through
    this proxy: it makes the device of the stream argument current for the duration of the call and refuses
    tensors that live on different devices (the library would fail with "invalid resource handle" or, worse,
    touch the wrong memory)."""

    def __getattr__(self, name):
        fn = getattr(_lib.lib(), name)

        def call(*args):
            dev = getattr(args[-1], "device", None) if args else None
            if dev is None or dev.type != "cuda":
                return fn(*args)
            for a in args:
                d = getattr(a, "device", None)
                if d is not None and d != dev:
                    raise RuntimeError("%s: tensors on different devices (%s vs %s)" % (name, d, dev))
            with torch.cuda.device(dev):
                return fn(*args)

        return call


_guarded = _Guarded()


def lib():
    return _guarded


def _require(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda and not _lib.host_pointers():
        raise RuntimeError("%s must be a CUDA tensor: npde-b200 has no CPU path" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32 (got %s)" % (name, t.dtype))
    return t.detach().contiguous()


def _opt(t, name):
    return None if t is None else _require(t, name)


def _ptr(t):
    if t is None:
        return None
    p = _Ptr(t.data_ptr())
    p.device = t.device
    return p


def _stream(t):
    """The current stream of t's device, tagged with that device (see _Guarded); None on the emulator."""
    if t.is_cuda:
        p = _Ptr(torch.cuda.current_stream(t.device).cuda_stream)
        p.device = t.device
        return p
    return None


def _workspace(nbytes, like):
    """Scratch for one call: cached per (device, stream), grown on demand, 256-byte aligned."""
    if nbytes == 0:
        return None, 0
    if like.is_cuda:
        key = (like.device, torch.cuda.current_stream(like.device).cuda_stream)
    else:
        key = (like.device, 0)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes + 256:
        buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=like.device)
        _workspaces[key] = buf
    base = buf.data_ptr()
    off = (-base) % 256
    return ctypes.c_void_p(base + off), buf.numel() - off


def bc_kind(x, bc):
    """utils/heat_utils.py:29-42 is_bc_mask, as the C enum.  Raises like the reference on a bad shape."""
    B, N = x.shape[0], x.shape[-1]
    if tuple(bc.shape) == (B, 2, N, N):
        return BC_MASK
    if tuple(bc.shape) == (B, 4):
        return BC_SQUARE
    print(x.shape, bc.shape)
    raise Exception("bc must be (B,4) or (B,2,N,N)")


def _field(x, name="x"):
    x = _require(x, name)
    if x.dim() != 3 or x.shape[1] != x.shape[2]:
        raise ValueError("%s must be (B, N, N), got %s" % (name, tuple(x.shape)))
    return x


def _act(act):
    if act not in ACT:
        raise NotImplementedError("activation %r" % (act,))  # models/iterators/iterator.py:25-26
    return ACT[act]


# ----------------------------------------------------------------- grid ops --
def set_boundary_(x, bc):
    """G in place (utils/heat_utils.py:44-59)."""
    xc = _field(x)
    if xc.data_ptr() != x.data_ptr():
        raise ValueError("set_boundary_ needs a contiguous tensor")
    bc = _require(bc, "bc")
    B, N, _ = xc.shape
    check(lib().npde_set_boundary(_ptr(xc), _ptr(bc), bc_kind(xc, bc), B, N, _stream(xc)), "npde_set_boundary")
    return x


def initialize_(x, bc, mode, rnd=None):
    """utils/heat_utils.py:70-94 in place on a contiguous (B,N,N) tensor; rnd: the uniform numbers for 'random'."""
    xc = _field(x)
    if xc.data_ptr() != x.data_ptr():
        raise ValueError("initialize_ needs a contiguous tensor")
    bc = _require(bc, "bc")
    rnd = _opt(rnd, "rnd")
    B, N, _ = xc.shape
    check(lib().npde_initialize(_ptr(xc), _ptr(bc), bc_kind(xc, bc), _lib.INIT[mode], _ptr(rnd), B, N, _stream(xc)),
          "npde_initialize")
    return x


def pad_boundary(xin, bc):
    xin = _field(xin, "xin")
    bc = _require(bc, "bc")
    B, S, _ = xin.shape
    N = S + 2
    y = torch.empty((B, N, N), dtype=torch.float32, device=xin.device)
    check(lib().npde_pad_boundary(_ptr(xin), _ptr(bc), bc_kind(y, bc), _ptr(y), B, N, _stream(xin)),
          "npde_pad_boundary")
    return y


def fd_step(x, bc, f=None):
    x = _field(x)
    bc = _require(bc, "bc")
    f = _opt(f, "f")
    B, N, _ = x.shape
    y = torch.empty_like(x)
    check(lib().npde_fd_step(_ptr(x), _ptr(bc), bc_kind(x, bc), _ptr(f), _ptr(y), B, N, _stream(x)), "npde_fd_step")
    return y


def fd_error(x, bc, f=None, aggregate="max"):
    if aggregate not in AGG:
        raise NotImplementedError(aggregate)  # heat_utils.py:130-131
    x = _field(x)
    bc = _require(bc, "bc")
    f = _opt(f, "f")
    B, N, _ = x.shape
    err = torch.empty((B,), dtype=torch.float32, device=x.device)
    kind = bc_kind(x, bc)
    ws, wsb = _workspace(lib().npde_workspace_bytes(_lib.WS_REDUCE, B, N, kind, 0, 0, 0, 0), x)
    check(lib().npde_fd_error(_ptr(x), _ptr(bc), kind, _ptr(f), AGG[aggregate], _ptr(err), B, N, ws, wsb,
                              _stream(x)), "npde_fd_error")
    return err


def l2_error(x, gt):
    if x.dim() != 3:
        raise NotImplementedError  # heat_utils.py:140-141
    x = _field(x)
    gt = _field(gt, "gt")
    B, N, _ = x.shape
    err = torch.empty((B,), dtype=torch.float32, device=x.device)
    ws, wsb = _workspace(lib().npde_workspace_bytes(_lib.WS_REDUCE, B, N, 0, 0, 0, 0, 0), x)
    check(lib().npde_l2_error(_ptr(x), _ptr(gt), _ptr(err), B, N, ws, wsb, _stream(x)), "npde_l2_error")
    return err


def subsample(x):
    x = _field(x)
    B, N, _ = x.shape
    M = (N - 1) // 2 + 1
    y = torch.empty((B, M, M), dtype=torch.float32, device=x.device)
    check(lib().npde_subsample(_ptr(x), _ptr(y), B, N, _stream(x)), "npde_subsample")
    return y


def restriction(x, bc_sub):
    x = _field(x)
    bc_sub = _require(bc_sub, "bc")
    B, N, _ = x.shape
    M = (N - 1) // 2 + 1
    y = torch.empty((B, M, M), dtype=torch.float32, device=x.device)
    check(lib().npde_restrict(_ptr(x), _ptr(bc_sub), bc_kind(y, bc_sub), _ptr(y), B, N, _stream(x)), "npde_restrict")
    return y


def interpolation(x, bc):
    x = _field(x)
    bc = _require(bc, "bc")
    B, M, _ = x.shape
    N = 2 * M - 1
    y = torch.empty((B, N, N), dtype=torch.float32, device=x.device)
    check(lib().npde_prolong(_ptr(x), _ptr(bc), bc_kind(y, bc), _ptr(y), B, M, _stream(x)), "npde_prolong")
    return y


# ------------------------------------------------------------------- Conv --
def _weights(w, name="w"):
    w = _require(w, name)
    return w.reshape(-1, 3, 3)


def conv_H(r, bc, w, is_bc_mask):
    """ConvIterator.H on r (B,N-2,N-2) (conv_iterator.py:19-31)."""
    r = _field(r, "r")
    w = _weights(w)
    B, S, _ = r.shape
    N = S + 2
    kind = BC_MASK if is_bc_mask else BC_SQUARE
    bcp = _require(bc, "bc") if is_bc_mask else None
    out = torch.empty_like(r)
    L = int(w.shape[0])
    ws, wsb = _workspace(lib().npde_workspace_bytes(_lib.WS_CONV_FWD, B, N, kind, L, 0, 0, 0), r)
    check(lib().npde_conv_H(_ptr(r), _ptr(bcp), kind, _ptr(w), L, _ptr(out), B, N, ws, wsb, _stream(r)),
          "npde_conv_H")
    return out


def conv_step(x, bc, f, w, act, w_host=None):
    """ConvIterator.forward without autograd.  w: (L,3,3) device; w_host: numpy copy (enables the fused kernel)."""
    x = _field(x)
    bc = _require(bc, "bc")
    f = _opt(f, "f")
    w = _weights(w)
    B, N, _ = x.shape
    L = int(w.shape[0])
    kind = bc_kind(x, bc)
    y = torch.empty_like(x)
    wh = None
    if w_host is not None:
        w_host = np.ascontiguousarray(w_host, dtype=np.float32).reshape(-1)
        assert w_host.size == 9 * L
        wh = w_host.ctypes.data_as(ctypes.c_void_p)
    ws, wsb = _workspace(lib().npde_workspace_bytes(_lib.WS_CONV_FWD, B, N, kind, L, 0, 0, 0), x)
    check(lib().npde_conv_step_fwd(_ptr(x), _ptr(bc), kind, _ptr(f), _ptr(w), wh, L, _act(act), _ptr(y), B, N, ws,
                                   wsb, _stream(x)), "npde_conv_step_fwd")
    return y


def conv_step_bwd(x, bc, f, w, act, gout, need_gx):
    x = _field(x)
    bc = _require(bc, "bc")
    f = _opt(f, "f")
    w = _weights(w)
    gout = _field(gout, "gout")
    B, N, _ = x.shape
    L = int(w.shape[0])
    kind = bc_kind(x, bc)
    gw = torch.empty((L, 3, 3), dtype=torch.float32, device=x.device)
    gx = torch.empty_like(x) if need_gx else None
    ws, wsb = _workspace(lib().npde_workspace_bytes(_lib.WS_CONV_BWD, B, N, kind, L, 0, 0, 0), x)
    check(lib().npde_conv_step_bwd(_ptr(x), _ptr(bc), kind, _ptr(f), _ptr(w), L, _act(act), _ptr(gout), _ptr(gw),
                                   _ptr(gx), B, N, ws, wsb, _stream(x)), "npde_conv_step_bwd")
    return gw, gx


class ConvStepFn(torch.autograd.Function):
    """Differentiable ConvIterator.forward: the reference gets this from autograd over
    F.conv2d (models/heat_model.py:52-53,71-72); here both directions are C-ABI calls."""

    @staticmethod
    def forward(ctx, x, bc, f, w, act, w_host):
        ctx.save_for_backward(x, bc, f if f is not None else x.new_empty(0), w)
        ctx.has_f = f is not None
        ctx.act = act
        return conv_step(x, bc, f, w, act, w_host)

    @staticmethod
    def backward(ctx, gout):
        x, bc, f, w = ctx.saved_tensors
        gw, gx = conv_step_bwd(x, bc, f if ctx.has_f else None, w, ctx.act, gout, ctx.needs_input_grad[0])
        return gx, None, None, gw.reshape(w.shape), None, None


class ConvUnrollFn(torch.autograd.Function):
    """k differentiable ConvIterator steps as ONE autograd node: forward keeps the k iterates
    (npde_conv_iterate_tape), backward walks them back in a single library call (npde_conv_iterate_bwd).
    Same gradients as k chained ConvStepFn nodes, without k round trips through the autograd engine."""

    @staticmethod
    def forward(ctx, x, bc, f, w, act, w_host, k):
        x = _field(x)
        bc = _require(bc, "bc")
        f = _opt(f, "f")
        wd = _weights(w)
        B, N, _ = x.shape
        L = int(wd.shape[0])
        kind = bc_kind(x, bc)
        tape = torch.empty((k, B, N, N), dtype=torch.float32, device=x.device)
        y = torch.empty_like(x)
        ws, wsb = _workspace(lib().npde_workspace_bytes(_lib.WS_CONV_FWD, B, N, kind, L, 0, 0, 0), x)
        w_host_c = None if w_host is None else np.ascontiguousarray(w_host, np.float32)  # stays alive over the call
        wh = None if w_host_c is None else ctypes.c_void_p(w_host_c.ctypes.data)
        check(lib().npde_conv_iterate_tape(_ptr(x), _ptr(bc), kind, _ptr(f), _ptr(wd), wh, L, _act(act), int(k),
                                           _ptr(tape), _ptr(y), B, N, ws, wsb, _stream(x)), "npde_conv_iterate_tape")
        ctx.save_for_backward(tape, bc, f if f is not None else x.new_empty(0), wd)
        ctx.has_f, ctx.act, ctx.k, ctx.w_shape = f is not None, act, int(k), w.shape
        return y

    @staticmethod
    def backward(ctx, gout):
        tape, bc, f, w = ctx.saved_tensors
        f = f if ctx.has_f else None
        gout = _field(gout, "gout")
        _, B, N, _ = tape.shape
        L = int(w.shape[0])
        kind = bc_kind(tape[0], bc)
        gw = torch.empty((L, 3, 3), dtype=torch.float32, device=tape.device)
        gx = torch.empty_like(gout) if ctx.needs_input_grad[0] else None
        ws, wsb = _workspace(lib().npde_workspace_bytes(_lib.WS_CONV_BPTT, B, N, kind, L, 0, 0, 0), tape)
        check(lib().npde_conv_iterate_bwd(_ptr(tape), _ptr(bc), kind, _ptr(f), _ptr(w), L, _act(ctx.act), ctx.k,
                                          _ptr(gout), _ptr(gw), _ptr(gx), B, N, ws, wsb, _stream(tape)),
              "npde_conv_iterate_bwd")
        return gx, None, None, gw.reshape(ctx.w_shape), None, None, None


# -------------------------------------------------------- Multigrid / UNet --
def cg_step(x, n_iters):
    """ConjugateGradient.forward (conjugate_gradient.py:15-47): n_iters CG iterations, boundary ring kept."""
    x = _field(x)
    B, N, _ = x.shape
    y = torch.empty_like(x)
    ws, wsb = _workspace(lib().npde_workspace_bytes(_lib.WS_CG, B, N, 0, 0, 0, 0, 0), x)
    check(lib().npde_cg_step(_ptr(x), int(n_iters), _ptr(y), B, N, ws, wsb, _stream(x)), "npde_cg_step")
    return y


def mg_step(x, bc, f, n_layers, pre, post):
    x = _field(x)
    bc = _require(bc, "bc")
    f = _opt(f, "f")
    B, N, _ = x.shape
    kind = bc_kind(x, bc)
    y = torch.empty_like(x)
    ws, wsb = _workspace(lib().npde_workspace_bytes(_lib.WS_MG, B, N, kind, n_layers, pre, post, 0), x)
    check(lib().npde_mg_step(_ptr(x), _ptr(bc), kind, _ptr(f), n_layers, pre, post, _ptr(y), B, N, ws, wsb,
                             _stream(x)), "npde_mg_step")
    return y


def unet_H(r, bc, w_first, w_pool, w_second, n_layers, pre, post, is_bc_mask):
    r = _field(r, "r")
    B, S, _ = r.shape
    N = S + 2
    kind = BC_MASK if is_bc_mask else BC_SQUARE
    bcp = _require(bc, "bc") if is_bc_mask else None
    wf, wp, wsn = _opt(w_first, "w_first"), _opt(w_pool, "w_pool"), _opt(w_second, "w_second")
    out = torch.empty_like(r)
    ws, wsb = _workspace(lib().npde_workspace_bytes(_lib.WS_UNET_FWD, B, N, kind, n_layers, pre, post, 0), r)
    check(lib().npde_unet_H(_ptr(r), _ptr(bcp), kind, _ptr(wf), _ptr(wp), _ptr(wsn), n_layers, pre, post, _ptr(out),
                            B, N, ws, wsb, _stream(r)), "npde_unet_H")
    return out


def unet_step(x, bc, f, w_first, w_pool, w_second, n_layers, pre, post, act, w_host=None):
    """w_host: optional host (numpy, float32) copy of the weights, first | pool | second concatenated: enables the
    fused finest level (npde_unet_step_fwd_hw)."""
    x = _field(x)
    bc = _require(bc, "bc")
    f = _opt(f, "f")
    wf, wp, wsn = _opt(w_first, "w_first"), _opt(w_pool, "w_pool"), _opt(w_second, "w_second")
    B, N, _ = x.shape
    kind = bc_kind(x, bc)
    y = torch.empty_like(x)
    ws, wsb = _workspace(lib().npde_workspace_bytes(_lib.WS_UNET_FWD, B, N, kind, n_layers, pre, post, 0), x)
    wh = None
    if w_host is not None:
        w_host = np.ascontiguousarray(w_host, dtype=np.float32)
        wh = ctypes.c_void_p(w_host.ctypes.data)
    check(lib().npde_unet_step_fwd_hw(_ptr(x), _ptr(bc), kind, _ptr(f), _ptr(wf), _ptr(wp), _ptr(wsn), wh, n_layers,
                                      pre, post, _act(act), _ptr(y), B, N, ws, wsb, _stream(x)),
          "npde_unet_step_fwd_hw")
    return y


def unet_step_bwd(x, bc, f, w_first, w_pool, w_second, n_layers, pre, post, act, gout, need_gx):
    x = _field(x)
    bc = _require(bc, "bc")
    f = _opt(f, "f")
    wf, wp, wsn = _opt(w_first, "w_first"), _opt(w_pool, "w_pool"), _opt(w_second, "w_second")
    gout = _field(gout, "gout")
    B, N, _ = x.shape
    kind = bc_kind(x, bc)
    gwf, gwp, gws = torch.zeros_like(wf), torch.zeros_like(wp), torch.zeros_like(wsn)
    gx = torch.empty_like(x) if need_gx else None
    ws, wsb = _workspace(lib().npde_workspace_bytes(_lib.WS_UNET_BWD, B, N, kind, n_layers, pre, post, 0), x)
    check(lib().npde_unet_step_bwd(_ptr(x), _ptr(bc), kind, _ptr(f), _ptr(wf), _ptr(wp), _ptr(wsn), n_layers, pre,
                                   post, _act(act), _ptr(gout), _ptr(gwf), _ptr(gwp), _ptr(gws), _ptr(gx), B, N, ws,
                                   wsb, _stream(x)), "npde_unet_step_bwd")
    return gwf, gwp, gws, gx


class UNetStepFn(torch.autograd.Function):
    """Differentiable UNetIterator.forward (the reference gets it from autograd, models/heat_model.py:52-53)."""

    @staticmethod
    def forward(ctx, x, bc, f, w, n_first, n_pool, n_layers, pre, post, act):
        ctx.save_for_backward(x, bc, f if f is not None else x.new_empty(0), w)
        ctx.has_f = f is not None
        ctx.cfg = (n_first, n_pool, n_layers, pre, post, act)
        wd = w.detach()
        return unet_step(x, bc, f, wd[:n_first], wd[n_first:n_first + n_pool], wd[n_first + n_pool:], n_layers, pre,
                         post, act, w_host=wd.cpu().numpy())

    @staticmethod
    def backward(ctx, gout):
        x, bc, f, w = ctx.saved_tensors
        n_first, n_pool, n_layers, pre, post, act = ctx.cfg
        wd = w.detach()
        gwf, gwp, gws, gx = unet_step_bwd(x, bc, f if ctx.has_f else None, wd[:n_first],
                                          wd[n_first:n_first + n_pool], wd[n_first + n_pool:], n_layers, pre, post,
                                          act, gout, ctx.needs_input_grad[0])
        return gx, None, None, torch.cat([gwf, gwp, gws]), None, None, None, None, None, None


# -------------------------------------------------------------- k-step loop --
def iterate(iterator, x, bc, f, k, *, w=None, w_host=None, n_layers=0, pre=0, post=0, act="none", gt=None,
            want_l2=False, want_fd=False, temporal_depth=0, out=None):
    """k applications of one iterator in a single C call (npde_iterate).

    iterator: _lib.IT_*.  Returns (x_k, err_l2 or None, err_fd or None) with
    err_l2 (k,B) = l2_error(x_{t+1}, gt) and err_fd (k+1,B) = fd_error(x_t,'max').
    """
    x = _field(x)
    bc = _require(bc, "bc")
    f = _opt(f, "f")
    gt = _opt(gt, "gt")
    w = None if w is None else _require(w, "w").reshape(-1)
    B, N, _ = x.shape
    kind = bc_kind(x, bc)
    for name, t in (("bc", bc), ("f", f), ("gt", gt), ("w", w)):
        if t is not None and t.device != x.device:
            raise RuntimeError("iterate: %s is on %s, x on %s" % (name, t.device, x.device))
    if out is not None:  # its pointer goes to the library as is: anything else would be a silent stray write
        if not isinstance(out, torch.Tensor) or out.dtype != torch.float32 or out.shape != x.shape or \
                out.device != x.device or not out.is_contiguous():
            raise ValueError("iterate: out must be a contiguous float32 tensor with x's shape and device")
    y = torch.empty_like(x) if out is None else out
    e2 = torch.empty((k, B), dtype=torch.float32, device=x.device) if want_l2 else None
    efd = torch.empty((k + 1, B), dtype=torch.float32, device=x.device) if want_fd else None
    if want_l2 and gt is None:
        raise ValueError("want_l2 needs gt")
    wh = None
    if w_host is not None:
        w_host = np.ascontiguousarray(w_host, dtype=np.float32).reshape(-1)
        wh = w_host.ctypes.data_as(ctypes.c_void_p)
    ws, wsb = _workspace(lib().npde_workspace_bytes(_lib.WS_ITERATE + iterator, B, N, kind, n_layers, pre, post, k), x)
    d = IterateDesc(iterator=iterator, B=B, N=N, bc_kind=kind, n_layers=n_layers, pre=pre, post=post, act=_act(act),
                    k=k, temporal_depth=temporal_depth, x_in=x.data_ptr(), x_out=y.data_ptr(), bc=bc.data_ptr(),
                    f=None if f is None else f.data_ptr(), w=None if w is None else w.data_ptr(), w_host=wh,
                    gt=None if gt is None else gt.data_ptr(), err_l2=None if e2 is None else e2.data_ptr(),
                    err_fd=None if efd is None else efd.data_ptr(), workspace=ws, workspace_bytes=wsb)
    check(lib().npde_iterate(ctypes.byref(d), _stream(x)), "npde_iterate")
    return y, e2, efd


def solve(iterator, x, bc, f, tol, check_every, max_iters, *, w=None, w_host=None, n_layers=0, pre=0, post=0,
          act="none"):
    """generation.py:72-99 as one device-side loop (npde_solve): apply the iterator until max_b fd_error < tol,
    checked every `check_every` applications, at most `max_iters`.  Only enqueues: returns (x_final, result) with
    result a 4-element int32 DEVICE tensor [iterations, converged, bits(residual), bits(residual0)] that is valid once
    the stream has run (read it with .cpu(): that is the only synchronisation)."""
    x = _field(x)
    bc = _require(bc, "bc")
    f = _opt(f, "f")
    w = None if w is None else _require(w, "w").reshape(-1)
    for name, t in (("bc", bc), ("f", f), ("w", w)):
        if t is not None and t.device != x.device:
            raise RuntimeError("solve: %s is on %s, x on %s" % (name, t.device, x.device))
    B, N, _ = x.shape
    kind = bc_kind(x, bc)
    y = torch.empty_like(x)
    result = torch.zeros(4, dtype=torch.int32, device=x.device)
    wh = None
    if w_host is not None:
        w_host = np.ascontiguousarray(w_host, dtype=np.float32).reshape(-1)
        wh = w_host.ctypes.data_as(ctypes.c_void_p)
    ws, wsb = _workspace(lib().npde_workspace_bytes(_lib.WS_ITERATE + iterator, B, N, kind, n_layers, pre, post,
                                                    int(check_every)), x)
    d = IterateDesc(iterator=iterator, B=B, N=N, bc_kind=kind, n_layers=n_layers, pre=pre, post=post, act=_act(act),
                    k=int(check_every), temporal_depth=0, x_in=x.data_ptr(), x_out=y.data_ptr(), bc=bc.data_ptr(),
                    f=None if f is None else f.data_ptr(), w=None if w is None else w.data_ptr(), w_host=wh,
                    gt=None, err_l2=None, err_fd=None, workspace=ws, workspace_bytes=wsb)
    check(lib().npde_solve(ctypes.byref(d), ctypes.c_float(tol), int(check_every), int(max_iters), _ptr(result),
                           _stream(x)), "npde_solve")
    return y, result


# ---------------------------------------------------------------- packed (W8) fields: npde.h "persistent layout" --
class PackedField(object):
    """A batch of square-domain fields in the batch-interleaved W8 layout of the fused Phi_H kernel (npde.h).
    A solver loop that packs its iterate once, runs `conv_iterate_packed` / `fd_error_packed` on it and unpacks the
    solution at the end pays no layout pass per call (drop-in callers of `iterate` pay two per call)."""

    def __init__(self, data, B, N):
        self.data, self.B, self.N = data, int(B), int(N)

    @property
    def device(self):
        return self.data.device

    def empty_like(self):
        return PackedField(torch.zeros_like(self.data), self.B, self.N)


def pack(x):
    """(B,N,N) -> PackedField (npde_pack)."""
    x = _require(x, "x")
    B, N = x.shape[0], x.shape[-1]
    data = torch.empty(int(lib().npde_packed_floats(B, N)), dtype=torch.float32, device=x.device)
    check(lib().npde_pack(_ptr(x), _ptr(data), B, N, _stream(x)), "npde_pack")
    return PackedField(data, B, N)


def unpack(p, out=None):
    """PackedField -> (B,N,N) (npde_unpack)."""
    if out is None:
        out = torch.empty((p.B, p.N, p.N), dtype=torch.float32, device=p.device)
    elif out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != (p.B, p.N, p.N) or out.device != p.device:
        raise ValueError("out must be a contiguous float32 (B,N,N) tensor on the field's device")
    check(lib().npde_unpack(_ptr(p.data), _ptr(out), p.B, p.N, _stream(out)), "npde_unpack")
    return out


def conv_iterate_packed(a, b, bc, f_packed, w_host, n_layers, act, k):
    """k applications of Phi_H ping-ponging between the PackedFields a (input) and b (scratch); returns the one that
    holds the k-th iterate.  w_host: (n_layers, 9) float32 numpy array (the weights travel as kernel parameters)."""
    bc = _require(bc, "bc")
    w = np.ascontiguousarray(w_host, dtype=np.float32).reshape(-1)
    if w.size != 9 * n_layers:
        raise ValueError("w_host must hold n_layers * 9 floats")
    fp = None if f_packed is None else _ptr(f_packed.data)
    check(lib().npde_conv_iterate_packed(_ptr(a.data), _ptr(b.data), _ptr(bc), fp, ctypes.c_void_p(w.ctypes.data), n_layers,
                                         ACT[act], a.B, a.N, int(k), _stream(a.data)), "npde_conv_iterate_packed")
    return a if k % 2 == 0 else b


def fd_error_packed(p, f_packed=None):
    """fd_error(x, bc, f, aggregate='max') of a PackedField -> (B,) (npde_fd_error_packed)."""
    err = torch.empty(p.B, dtype=torch.float32, device=p.device)
    fp = None if f_packed is None else _ptr(f_packed.data)
    check(lib().npde_fd_error_packed(_ptr(p.data), fp, _ptr(err), p.B, p.N, _stream(err)), "npde_fd_error_packed")
    return err
