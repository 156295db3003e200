"""CPU oracle bindings (TEST INFRASTRUCTURE -- never imported by the product path).

ctypes wrappers over ``oracle/_build/libnpde_oracle.so`` (``oracle/npde_oracle.c``,
a plain-C restatement of the reference hot path).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
All functions take / return C-contiguous float32 numpy arrays in the
reference's layouts.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnpde_oracle.so")

BC_SQUARE, BC_MASK = 0, 1
ACT = {"none": 0, "clamp": 1, "sigmoid": 2}
AGG = {"max": 0, "mean": 1}


def build(force=False):
    """Compile the C oracle with the recipe in oracle/Makefile."""
    src = os.path.join(_HERE, "npde_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], env=env)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def set_threads(n=0):
    """Use n OpenMP threads for the batch loops (n <= 0: only query).  Returns the thread count in effect."""
    fn = lib().orc_set_threads
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int]
    return int(fn(int(n)))


def _f(a):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def bc_kind(x_shape, bc):
    """utils/heat_utils.py:29-42 is_bc_mask."""
    B, N = x_shape[0], x_shape[-1]
    if bc.shape == (B, 2, N, N):
        return BC_MASK
    if bc.shape == (B, 4):
        return BC_SQUARE
    raise Exception("bad bc shape %r for x %r" % (bc.shape, x_shape))


def set_boundary(x, bc):
    x = np.array(x, dtype=np.float32, order="C", copy=True)
    bc_, pbc = _f(bc)
    B, N, _ = x.shape
    lib().orc_set_boundary(x.ctypes.data_as(ctypes.c_void_p), pbc, bc_kind(x.shape, bc_), B, N)
    return x


def pad_boundary(xin, bc):
    xin_, px = _f(xin)
    bc_, pbc = _f(bc)
    B, S, _ = xin_.shape
    N = S + 2
    y = np.empty((B, N, N), np.float32)
    lib().orc_pad_boundary(px, pbc, bc_kind(y.shape, bc_), y.ctypes.data_as(ctypes.c_void_p), B, N)
    return y


def fd_step(x, bc, f=None):
    x_, px = _f(x)
    bc_, pbc = _f(bc)
    f_, pf = _f(f)
    B, N, _ = x_.shape
    y = np.empty_like(x_)
    lib().orc_fd_step(px, pbc, bc_kind(x_.shape, bc_), pf, y.ctypes.data_as(ctypes.c_void_p), B, N)
    return y


def fd_error(x, bc, f=None, aggregate="max"):
    x_, px = _f(x)
    bc_, pbc = _f(bc)
    f_, pf = _f(f)
    B, N, _ = x_.shape
    e = np.empty((B,), np.float32)
    lib().orc_fd_error(px, pbc, bc_kind(x_.shape, bc_), pf, AGG[aggregate],
                       e.ctypes.data_as(ctypes.c_void_p), B, N)
    return e


def l2_error(x, gt):
    x_, px = _f(x)
    g_, pg = _f(gt)
    B, N, _ = x_.shape
    e = np.empty((B,), np.float32)
    lib().orc_l2_error(px, pg, e.ctypes.data_as(ctypes.c_void_p), B, N)
    return e


def conv_H(r, bc, w, is_bc_mask):
    """r: (B, N-2, N-2); w: (L,3,3).  bc only used when is_bc_mask."""
    r_, pr = _f(r)
    w_, pw = _f(w)
    B, S, _ = r_.shape
    N = S + 2
    out = np.empty_like(r_)
    if is_bc_mask:
        bc_, pbc = _f(bc)
        kind = BC_MASK
    else:
        pbc, kind = None, BC_SQUARE
    lib().orc_conv_H(pr, pbc, kind, pw, int(w_.shape[0]), out.ctypes.data_as(ctypes.c_void_p), B, N)
    return out


def conv_step(x, bc, f, w, act):
    x_, px = _f(x)
    bc_, pbc = _f(bc)
    f_, pf = _f(f)
    w_, pw = _f(w)
    B, N, _ = x_.shape
    y = np.empty_like(x_)
    lib().orc_conv_step(px, pbc, bc_kind(x_.shape, bc_), pf, pw, int(w_.shape[0]), ACT[act],
                        y.ctypes.data_as(ctypes.c_void_p), B, N)
    return y


def subsample(x):
    x_, px = _f(x)
    B, N, _ = x_.shape
    M = (N - 1) // 2 + 1
    y = np.empty((B, M, M), np.float32)
    lib().orc_subsample(px, y.ctypes.data_as(ctypes.c_void_p), B, N)
    return y


def restriction(x, bc_sub):
    x_, px = _f(x)
    bc_, pbc = _f(bc_sub)
    B, N, _ = x_.shape
    M = (N - 1) // 2 + 1
    y = np.empty((B, M, M), np.float32)
    lib().orc_restriction(px, pbc, bc_kind(y.shape, bc_), y.ctypes.data_as(ctypes.c_void_p), B, N)
    return y


def interpolation(x, bc):
    x_, px = _f(x)
    bc_, pbc = _f(bc)
    B, M, _ = x_.shape
    N = 2 * M - 1
    y = np.empty((B, N, N), np.float32)
    lib().orc_interpolation(px, pbc, bc_kind(y.shape, bc_), y.ctypes.data_as(ctypes.c_void_p), B, M)
    return y


def multigrid_step(x, bc, f, n_layers, pre, post):
    x_, px = _f(x)
    bc_, pbc = _f(bc)
    f_, pf = _f(f)
    B, N, _ = x_.shape
    y = np.empty_like(x_)
    lib().orc_multigrid_step(px, pbc, bc_kind(x_.shape, bc_), pf, n_layers, pre, post,
                             y.ctypes.data_as(ctypes.c_void_p), B, N)
    return y


def unet_H(r, bc, w_first, w_pool, w_second, n_layers, pre, post, is_bc_mask):
    r_, pr = _f(r)
    wf, pwf = _f(w_first)
    wp, pwp = _f(w_pool)
    ws, pws = _f(w_second)
    B, S, _ = r_.shape
    N = S + 2
    out = np.empty_like(r_)
    if is_bc_mask:
        bc_, pbc = _f(bc)
        kind = BC_MASK
    else:
        pbc, kind = None, BC_SQUARE
    lib().orc_unet_H(pr, pbc, kind, pwf, pwp, pws, n_layers, pre, post,
                     out.ctypes.data_as(ctypes.c_void_p), B, N)
    return out


def unet_step(x, bc, f, w_first, w_pool, w_second, n_layers, pre, post, act):
    x_, px = _f(x)
    bc_, pbc = _f(bc)
    f_, pf = _f(f)
    wf, pwf = _f(w_first)
    wp, pwp = _f(w_pool)
    ws, pws = _f(w_second)
    B, N, _ = x_.shape
    y = np.empty_like(x_)
    lib().orc_unet_step(px, pbc, bc_kind(x_.shape, bc_), pf, pwf, pwp, pws, n_layers, pre, post,
                        ACT[act], y.ctypes.data_as(ctypes.c_void_p), B, N)
    return y


def conv_step_bwd(x, bc, f, w, act, gout, want_gx=True):
    x_, px = _f(x)
    bc_, pbc = _f(bc)
    f_, pf = _f(f)
    w_, pw = _f(w)
    g_, pg = _f(gout)
    B, N, _ = x_.shape
    L = int(w_.shape[0])
    gw = np.empty((L, 3, 3), np.float32)
    gx = np.empty_like(x_) if want_gx else None
    lib().orc_conv_step_bwd(px, pbc, bc_kind(x_.shape, bc_), pf, pw, L, ACT[act], pg,
                            gw.ctypes.data_as(ctypes.c_void_p),
                            gx.ctypes.data_as(ctypes.c_void_p) if want_gx else None, B, N)
    return gw, gx


def cg_step(x, n_iters):
    """ConjugateGradient.forward (conjugate_gradient.py:15-47), f = None."""
    x_, px = _f(x)
    B, N, _ = x_.shape
    y = np.empty_like(x_)
    lib().orc_cg_step(px, int(n_iters), y.ctypes.data_as(ctypes.c_void_p), B, N)
    return y


def multigrid_residual_step(x, bc, f, n_layers, pre, post):
    """MultigridResidualIterator.forward (jacobi_iterator.py:81-143) composed from the operators above."""
    def level(x, bc, f, step):
        if step == 0:
            return None
        x = set_boundary(x, bc)
        for _ in range(pre):
            x = fd_step(x, bc, f)
        r = fd_step(x, bc, f) - x
        zeros = np.zeros((x.shape[0], 4), np.float32)
        r_sub = restriction(r, zeros)
        ek = level(r_sub, zeros, -r_sub, step - 1)
        if ek is not None:
            x = x + interpolation(ek, zeros)
        x = set_boundary(x, bc)
        for _ in range(post):
            x = fd_step(x, bc, f)
        return x
    return level(np.ascontiguousarray(x, np.float32), bc, f, n_layers)


def iterate(iterator, x, bc, f, w, act, k, gt=None, want_fd=False):
    """iterator: 'jacobi' | 'conv'.  Returns (x_k, err_l2 (k,B) | None, err_fd (k+1,B) | None)."""
    x_, px = _f(x)
    bc_, pbc = _f(bc)
    f_, pf = _f(f)
    w_, pw = _f(w)
    g_, pg = _f(gt)
    B, N, _ = x_.shape
    L = int(w_.shape[0]) if w_ is not None else 0
    y = np.empty_like(x_)
    e2 = np.empty((k, B), np.float32) if gt is not None else None
    ef = np.empty((k + 1, B), np.float32) if want_fd else None
    lib().orc_iterate(0 if iterator == "jacobi" else 1, px, y.ctypes.data_as(ctypes.c_void_p), pbc,
                      bc_kind(x_.shape, bc_), pf, pw, L, ACT[act], k, pg,
                      e2.ctypes.data_as(ctypes.c_void_p) if e2 is not None else None,
                      ef.ctypes.data_as(ctypes.c_void_p) if ef is not None else None, B, N)
    return y, e2, ef
