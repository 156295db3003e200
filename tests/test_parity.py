"""Parity of the CUDA path (through the Python drop-in API and the C ABI) with
the reference's golden vectors and with the CPU oracle.

Every test runs on two backends:
  * ``cuda`` (marked gpu): the product -- libnpde_b200.so on a B200;
  * ``emu``  (CPU): the SAME .cu sources compiled for the host against the SIMT
    emulator of tests/emu (test infrastructure), so kernel index logic and the
    host-side schedules are exercised in the CPU suite as well.

Bar: bit-exact against the golden vectors wherever the oracle is (see
tests/test_oracle_golden.py); otherwise <= 1e-5 relative (BASELINE.json
north_star), stated next to each assert.
"""
import importlib

import numpy as np
import pytest
import torch

import emu_backend
from conftest import assert_bitwise, golden_names, load_golden, rel_err

ACTS = {0: "none", 1: "clamp", 2: "sigmoid"}
BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


class Backend(object):
    def __init__(self, kind):
        self.kind = kind
        self.npde = importlib.import_module("neural-pde-solver_b200")
        if kind == "emu":
            emu_backend.use_emulator(self.npde)
            self.device = torch.device("cpu")
        else:
            assert torch.cuda.is_available(), "gpu tests need a CUDA device"
            emu_backend.use_cuda_library(self.npde)
            self.npde._lib.lib()  # raises if libnpde_b200.so is missing: no fallback
            self.device = torch.device("cuda:0")
        self.F = self.npde.functional
        self.utils = self.npde.utils

    def t(self, a):
        return None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)

    @staticmethod
    def n(t):
        return t.detach().cpu().numpy()


@pytest.fixture(params=BACKENDS)
def be(request):
    b = Backend(request.param)
    yield b
    emu_backend.use_cuda_library(b.npde)


# ------------------------------------------------------------ grid operators --
@pytest.mark.parametrize("name", golden_names("ops_"))
def test_grid_operators(be, name):
    g = load_golden(name)
    u = be.utils
    x, bc, gt = be.t(g["x"]), be.t(g["bc"]), be.t(g["gt"])
    assert_bitwise(be.n(u.fd_step(x, bc, None)), g["fd_step"], "fd_step")
    assert_bitwise(be.n(u.fd_error(x, bc, None, "max")), g["fd_error_max"], "fd_error max")
    assert rel_err(be.n(u.fd_error(x, bc, None, "mean")), g["fd_error_mean"]) < 1e-6
    f = be.t(g.get("f"))
    if f is not None:
        assert_bitwise(be.n(u.fd_step(x, bc, f)), g["fd_step_f"], "fd_step f")
        assert_bitwise(be.n(u.fd_error(x, bc, f, "max")), g["fd_error_max_f"], "fd_error f")
        assert rel_err(be.n(u.fd_error(x, bc, f, "mean")), g["fd_error_mean_f"]) < 1e-6
    assert rel_err(be.n(u.l2_error(x, gt)), g["l2_error"]) < 1e-6
    assert_bitwise(be.n(u.set_boundary(gt.clone(), bc)), g["set_boundary"], "set_boundary")
    assert_bitwise(be.n(u.pad_boundary(gt[:, 1:-1, 1:-1], bc)), g["pad_boundary"], "pad_boundary")
    assert_bitwise(be.n(u.subsample(x)), g["subsample"], "subsample")
    assert_bitwise(be.n(u.restriction(x, be.t(g["bc_sub"]))), g["restriction"], "restriction")
    assert_bitwise(be.n(u.interpolation(be.t(g["xs"]), bc)), g["interpolation"], "interpolation")
    # x is not modified by any of the calls above
    assert_bitwise(be.n(x), g["x"], "inputs untouched")


@pytest.mark.parametrize("name", golden_names("ops_"))
@pytest.mark.parametrize("depth", [0, 1, 2, 4])
def test_jacobi_chain(be, name, depth):
    """25 chained Jacobi steps (temporally blocked on the square path) + fd_error rows."""
    g = load_golden(name)
    x, bc, f = be.t(g["x"]), be.t(g["bc"]), be.t(g.get("f"))
    it = be.npde.JacobiIterator()
    y, _, _ = it.iterate(x, bc, f, 25, temporal_depth=depth)
    assert_bitwise(be.n(y), g["jacobi25"], "25 fused Jacobi steps, depth %d" % depth)
    if depth == 0:
        y2, _, efd = it.iterate(x, bc, f, 25, want_fd=True)
        assert_bitwise(be.n(y2), g["jacobi25"], "25 Jacobi steps with norms")
        assert_bitwise(be.n(efd[-1]), g["jacobi25_fd_error"], "fd_error row 25")
        assert_bitwise(be.n(efd[0]), g["fd_error_max_f" if f is not None else "fd_error_max"], "fd_error row 0")
        z = x
        for _ in range(25):
            z = it(z, bc, f)
        assert_bitwise(be.n(z), g["jacobi25"], "25 single Jacobi steps")


# ---------------------------------------------------------------- ConvIterator --
def make_conv(be, g):
    L, act, is_mask = (int(v) for v in g["meta"])
    it = be.npde.ConvIterator(ACTS[act], L)
    it.is_bc_mask = bool(is_mask)
    with torch.no_grad():
        for l, w in zip(it.layers, g["w"]):
            l.weight.copy_(torch.from_numpy(w).view(1, 1, 3, 3))
    return it.to(be.device)


@pytest.mark.parametrize("name", golden_names("conv_"))
def test_conv_iterator(be, name):
    g = load_golden(name)
    it = make_conv(be, g)
    x, bc, f = be.t(g["x"]), be.t(g["bc"]), be.t(g.get("f"))
    with torch.no_grad():
        y1 = it(x, bc, f)
        h = it.H(be.t(g["r"]).unsqueeze(1), bc)[:, 0]
        y10, _, _ = it.iterate(x, bc, f, 10)                       # <= 65^2: the one-kernel on-chip loop
        y10s, _, _ = it.iterate(x, bc, f, 10, temporal_depth=1)    # the streaming kernel, one launch per step
        z = x
        for _ in range(10):
            z = it.iter_step(z, bc, f)
    assert_bitwise(be.n(z), be.n(y10), "iterate(10) == 10 x forward")
    assert_bitwise(be.n(y10s), be.n(y10), "streaming loop == on-chip loop")
    if g["x"].shape[0] >= 2 and it.act != "sigmoid":
        assert_bitwise(be.n(h), g["H"], "H")
        assert_bitwise(be.n(y1), g["step1"], "conv step")
        assert_bitwise(be.n(y10), g["step10"], "10 conv steps")
    else:  # B == 1 changes ATen's conv order; sigmoid uses a different exp: north-star tolerance
        assert rel_err(be.n(h), g["H"]) <= 1e-5
        assert rel_err(be.n(y1), g["step1"]) <= 1e-5
        assert rel_err(be.n(y10), g["step10"]) <= 1e-5


def test_conv_zero_weights_is_jacobi(be):
    """ConvIterator with H = 0 (act none) equals JacobiIterator bitwise."""
    g = load_golden("ops_square33")
    x, bc, f = be.t(g["x"]), be.t(g["bc"]), be.t(g["f"])
    it = be.npde.ConvIterator("none", 3).to(be.device)
    with torch.no_grad():
        for l in it.layers:
            l.weight.zero_()
        y = it(x, bc, f)
    assert_bitwise(be.n(y), g["fd_step_f"], "zero-H conv == Jacobi")


def test_constant_field_is_fixed_point(be):
    """a constant field with equal bc is an exact fixed point of every iterator."""
    B, N = 2, 33
    x = torch.full((B, N, N), 0.375, device=be.device)
    bc = torch.full((B, 4), 0.375, device=be.device)
    its = [be.npde.JacobiIterator(), be.npde.ConvIterator("clamp", 3).to(be.device),
           be.npde.MultigridIterator(3, 2, 2), be.npde.UNetIterator("clamp", 3, 1, 1).to(be.device)]
    with torch.no_grad():
        for it in its:
            assert_bitwise(be.n(it(x, bc, None)), be.n(x), it.name())


# --------------------------------------------- work partition of the stream kernels --
@pytest.mark.parametrize("B,N,L,has_f", [(3, 129, 3, False), (5, 65, 2, True), (2, 257, 3, True), (3, 65, 4, False),
                                         (1, 257, 1, False), (7, 33, 3, True)])
def test_stream_partition_vs_oracle(be, oracle, B, N, L, has_f):
    """Odd batch sizes, several strips per image and several bands per strip (npde_stream.cu make_split).
    k-step loops (pair-permuted internal buffers, 128-bit accesses) and single steps (reference layout)
    against the CPU oracle, bit for bit."""
    rng = np.random.RandomState(B * 1000 + N + L)
    bc = rng.rand(B, 4).astype(np.float32)
    x = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    f = None
    if has_f:
        f = np.zeros((B, N, N), np.float32)
        f[:, 1:-1, 1:-1] = (rng.rand(B, N - 2, N - 2) - 0.5) * 1e-3
    w = ((rng.rand(L, 3, 3) - 0.5) * 0.3).astype(np.float32)
    it = be.npde.ConvIterator("clamp", L)
    with torch.no_grad():
        for l, wl in zip(it.layers, w):
            l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
    it = it.to(be.device)
    xd, bcd, fd = be.t(x), be.t(bc), be.t(f)
    with torch.no_grad():
        y1 = it(xd, bcd, fd)
        y5, _, _ = it.iterate(xd, bcd, fd, 5)
        j1 = be.npde.JacobiIterator()(xd, bcd, fd)
        j11, _, _ = be.npde.JacobiIterator().iterate(xd, bcd, fd, 11)
    assert_bitwise(be.n(y1), oracle.conv_step(x, bc, f, w, "clamp"), "conv step")
    assert_bitwise(be.n(y5), oracle.iterate("conv", x, bc, f, w, "clamp", 5)[0], "5 conv steps")
    assert_bitwise(be.n(j1), oracle.fd_step(x, bc, f), "jacobi step")
    assert_bitwise(be.n(j11), oracle.iterate("jacobi", x, bc, f, None, "none", 11)[0], "11 jacobi steps")


def _mask_geometry(rng, B, N):
    """Random obstacle geometries: ring + an L-shaped cut-out + discs, values in [0,1) on Dirichlet cells
    and 0 elsewhere (utils/geometries.py conventions: bc = [values, mask])."""
    m = np.zeros((B, N, N), np.float32)
    m[:, 0, :] = m[:, -1, :] = m[:, :, 0] = m[:, :, -1] = 1
    yy, xx = np.mgrid[0:N, 0:N]
    for b in range(B):
        if b % 2 == 0:
            m[b, : N // 2, N // 2:] = 1
        for _ in range(2):
            cy, cx = rng.randint(N // 5, 4 * N // 5, 2)
            r = rng.randint(max(N // 16, 1), max(N // 8, 2))
            m[b][(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 1
    v = (rng.rand(B, N, N).astype(np.float32)) * m
    return np.stack([v, m], 1)


@pytest.mark.parametrize("B,N,L,has_f,act", [(3, 129, 3, False, "clamp"), (2, 257, 3, True, "none"),
                                             (5, 65, 2, True, "clamp"), (3, 33, 4, False, "none"),
                                             (2, 129, 1, False, "clamp")])
def test_mask_geometry_stream_vs_oracle(be, oracle, B, N, L, has_f, act):
    """Fused Phi_H on mask geometries (bc = [values, mask]): single steps on the reference layout and
    k-step loops on the permuted buffers, several strips and bands per image, against the CPU oracle."""
    rng = np.random.RandomState(B * 77 + N + L)
    bc = _mask_geometry(rng, B, N)
    x = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    f = None
    if has_f:
        f = np.zeros((B, N, N), np.float32)
        f[:, 1:-1, 1:-1] = (rng.rand(B, N - 2, N - 2) - 0.5) * 1e-3
    w = ((rng.rand(L, 3, 3) - 0.5) * 0.3).astype(np.float32)
    it = be.npde.ConvIterator(act, L)
    it.is_bc_mask = True
    with torch.no_grad():
        for l, wl in zip(it.layers, w):
            l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
    it = it.to(be.device)
    xd, bcd, fd = be.t(x), be.t(bc), be.t(f)
    with torch.no_grad():
        y1 = it(xd, bcd, fd)
        y6, _, efd = it.iterate(xd, bcd, fd, 6, want_fd=True)
        y2, _, _ = it.iterate(xd, bcd, fd, 2)
    assert_bitwise(be.n(y1), oracle.conv_step(x, bc, f, w, act), "mask conv step")
    ref6, _, ref_fd = oracle.iterate("conv", x, bc, f, w, act, 6, want_fd=True)
    assert_bitwise(be.n(y6), ref6, "6 mask conv steps")
    assert_bitwise(be.n(efd), ref_fd, "fd_error rows on the mask geometry")
    assert_bitwise(be.n(y2), oracle.iterate("conv", x, bc, f, w, act, 2)[0], "2 mask conv steps (reference layout)")
    jac = be.npde.JacobiIterator()
    jac.is_bc_mask = True
    j7, _, jfd = jac.iterate(xd, bcd, fd, 7, want_fd=True)
    refj, _, refj_fd = oracle.iterate("jacobi", x, bc, f, None, "none", 7, want_fd=True)
    assert_bitwise(be.n(j7), refj, "7 Jacobi steps on the mask geometry (stream kernel)")
    assert_bitwise(be.n(jfd), refj_fd, "their fd_error rows")
    assert_bitwise(be.n(jac(xd, bcd, fd)), oracle.fd_step(x, bc, f), "single mask Jacobi step")


@pytest.mark.parametrize("B,N,L,pre,post,mask,has_f,act", [(3, 129, 3, 2, 2, True, False, "clamp"),
                                                            (2, 257, 2, 1, 3, False, True, "none"),
                                                            (5, 65, 3, 4, 1, True, True, "clamp"),
                                                            (2, 129, 1, 2, 2, False, False, "clamp"),
                                                            (3, 33, 2, 3, 4, True, False, "none"),
                                                            (17, 65, 3, 2, 2, True, False, "clamp"),
                                                            (16, 33, 4, 1, 2, False, True, "none"),
                                                            (16, 257, 3, 2, 2, False, False, "clamp")])
def test_unet_fused_level_vs_oracle(be, oracle, B, N, L, pre, post, mask, has_f, act):
    """UNetIterator.forward with the finest level fused into the head / tail stream kernels (several strips and
    bands, both geometries) against the CPU oracle; iterate() takes the same path.  From 16 problems on (always in
    the emulated build) every level of at most 127^2 runs in the on-chip kernel k_unet_coarse."""
    if be.kind == "emu" and B * N * N > 200000:
        pytest.skip("GPU-sized case")
    rng = np.random.RandomState(B * 31 + N + L + pre)
    bc = _mask_geometry(rng, B, N) if mask else rng.rand(B, 4).astype(np.float32)
    x = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    f = None
    if has_f:
        f = np.zeros((B, N, N), np.float32)
        f[:, 1:-1, 1:-1] = (rng.rand(B, N - 2, N - 2) - 0.5) * 1e-3
    wf = ((rng.rand(L * pre, 3, 3) - 0.5) * 0.3).astype(np.float32)
    wp = ((rng.rand(max(L - 1, 0), 3, 3) - 0.5) * 0.3).astype(np.float32)
    ws = ((rng.rand(L * post, 3, 3) - 0.5) * 0.3).astype(np.float32)
    it = be.npde.UNetIterator(act, L, pre, post)
    it.is_bc_mask = mask
    with torch.no_grad():
        for ml, wsrc in ((it.first_layers, wf), (it.pooling_layers, wp), (it.second_layers, ws)):
            for l, wl in zip(ml, wsrc):
                l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
    it = it.to(be.device)
    xd, bcd, fd = be.t(x), be.t(bc), be.t(f)
    with torch.no_grad():
        y1 = it(xd, bcd, fd)
        y3, _, _ = it.iterate(xd, bcd, fd, 3)
    ref = oracle.unet_step(x, bc, f, wf, wp, ws, L, pre, post, act)
    assert_bitwise(be.n(y1), ref, "unet step, fused finest level")
    for _ in range(2):
        ref = oracle.unet_step(ref, bc, f, wf, wp, ws, L, pre, post, act)
    assert_bitwise(be.n(y3), ref, "3 unet steps")


# ------------------------------------------------------------ Multigrid / UNet --
@pytest.mark.parametrize("name", golden_names("mg_"))
def test_multigrid_iterator(be, name):
    g = load_golden(name)
    nl, pre, post, is_mask = (int(v) for v in g["meta"])
    it = be.npde.MultigridIterator(nl, pre, post)
    it.is_bc_mask = bool(is_mask)
    x, bc, f = be.t(g["x"]), be.t(g["bc"]), be.t(g.get("f"))
    y1 = it(x, bc, f)
    assert_bitwise(be.n(y1), g["step1"], "V-cycle")
    y3, _, _ = it.iterate(x, bc, f, 3)
    assert_bitwise(be.n(y3), g["step3"], "3 V-cycles")


def make_unet(be, g):
    nl, pre, post, act, is_mask = (int(v) for v in g["meta"])
    it = be.npde.UNetIterator(ACTS[act], nl, pre, post)
    it.is_bc_mask = bool(is_mask)
    with torch.no_grad():
        for ml, ws in ((it.first_layers, g["w_first"]), (it.pooling_layers, g["w_pool"]),
                       (it.second_layers, g["w_second"])):
            for l, w in zip(ml, ws):
                l.weight.copy_(torch.from_numpy(w).view(1, 1, 3, 3))
    return it.to(be.device)


@pytest.mark.parametrize("name", golden_names("unet_"))
def test_unet_iterator(be, name):
    g = load_golden(name)
    it = make_unet(be, g)
    x, bc, f = be.t(g["x"]), be.t(g["bc"]), be.t(g.get("f"))
    with torch.no_grad():
        h = it.H(be.t(g["r"]).unsqueeze(1), bc)[:, 0]
        y1 = it(x, bc, f)
        y5, _, _ = it.iterate(x, bc, f, 5)
    assert_bitwise(be.n(h), g["H"], "unet H")
    assert_bitwise(be.n(y1), g["step1"], "unet step")
    assert_bitwise(be.n(y5), g["step5"], "5 unet steps")


# ------------------------------------------------------------------- backward --
@pytest.mark.parametrize("name", golden_names("bwd_conv_"))
def test_conv_backward(be, name):
    """One differentiable step + MSE, gradients vs the reference's autograd
    (models/heat_model.py:52-53,71-72).  Reduction order differs: tolerance 2e-5 of the largest entry."""
    g = load_golden(name)
    it = make_conv(be, g)
    x, bc, f, gt = be.t(g["x"]), be.t(g["bc"]), be.t(g.get("f")), be.t(g["gt"])
    xg = x.clone().requires_grad_(True)
    y = it(xg, bc, f)
    loss = torch.nn.MSELoss()(y, gt)
    loss.backward()
    assert rel_err(be.n(y), g["y"]) <= 1e-5
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    gw = np.stack([be.n(l.weight.grad)[0, 0] for l in it.layers])
    assert rel_err(gw, g["gw"]) <= 2e-5, (gw, g["gw"])
    assert rel_err(be.n(xg.grad), g["gx"]) <= 2e-5


@pytest.mark.parametrize("B,N,L,act,has_f", [(3, 65, 3, "clamp", False), (5, 33, 2, "none", True),
                                             (2, 129, 1, "sigmoid", False), (7, 17, 3, "clamp", True)])
def test_streaming_backward_vs_layered(be, B, N, L, act, has_f, monkeypatch):
    """k_bwd_stream (one fused pass: recompute + adjoint + weight gradient) against the layered backward kernels
    (NPDE_BWD_LAYERED=1), which the reference's autograd goldens pin: several strips and bands per image, strips
    that span images, an upstream gradient on every cell.  Same arithmetic per cell up to the summation order of
    the transposed convolution and of the weight-gradient sums: 1e-5 of the largest entry."""
    rng = np.random.RandomState(5 * B + N + L)
    bc = rng.rand(B, 4).astype(np.float32)
    x = rng.rand(B, N, N).astype(np.float32)
    f = None
    if has_f:
        f = np.zeros((B, N, N), np.float32)
        f[:, 1:-1, 1:-1] = (rng.rand(B, N - 2, N - 2) - 0.5) * 1e-2
    gout = (rng.rand(B, N, N).astype(np.float32) - 0.5)
    w = ((rng.rand(L, 3, 3) - 0.5) * 0.4).astype(np.float32)
    res = {}
    for mode in ("stream", "layered"):
        if mode == "layered":
            monkeypatch.setenv("NPDE_BWD_LAYERED", "1")
        else:
            monkeypatch.delenv("NPDE_BWD_LAYERED", raising=False)
        it = be.npde.ConvIterator(act, L)
        with torch.no_grad():
            for l, wl in zip(it.layers, w):
                l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
        it = it.to(be.device)
        xg = be.t(x).clone().requires_grad_(True)
        y = it(xg, be.t(bc), be.t(f))
        y.backward(be.t(gout))
        res[mode] = (be.n(xg.grad).copy(), np.stack([be.n(l.weight.grad)[0, 0] for l in it.layers]))
    assert rel_err(res["stream"][0], res["layered"][0]) <= 1e-5
    assert rel_err(res["stream"][1], res["layered"][1]) <= 1e-5, (res["stream"][1], res["layered"][1])


@pytest.mark.parametrize("name", golden_names("bwd_unet_"))
def test_unet_backward(be, name):
    """One differentiable UNetIterator step + MSE vs the reference's autograd (square and cylinder geometries,
    none / clamp / sigmoid).  Reduction order differs: 2e-5 of the largest entry per gradient tensor."""
    g = load_golden(name)
    it = make_unet(be, g)
    x, bc, f, gt = be.t(g["x"]), be.t(g["bc"]), be.t(g.get("f")), be.t(g["gt"])
    xg = x.clone().requires_grad_(True)
    y = it(xg, bc, f)
    loss = torch.nn.MSELoss()(y, gt)
    loss.backward()
    assert rel_err(be.n(y), g["y"]) <= 1e-5
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    for ml, key in ((it.first_layers, "gw_first"), (it.pooling_layers, "gw_pool"), (it.second_layers, "gw_second")):
        if len(ml) == 0:
            continue
        gw = np.stack([be.n(l.weight.grad)[0, 0] for l in ml])
        assert rel_err(gw, g[key]) <= 2e-5, (key, gw, g[key])
    assert rel_err(be.n(xg.grad), g["gx"]) <= 2e-5


@pytest.mark.parametrize("name", golden_names("bptt_"))
def test_backprop_through_k_steps(be, name):
    """BASELINE config 4: gradients through k unrolled iterations, vs the reference's iterator modules chained
    k times under autograd (no detach).  Tolerance 5e-5 of the largest entry (k reductions re-associated)."""
    g = load_golden(name)
    it = make_unet(be, g) if "w_first" in g else make_conv(be, g)
    x, bc, f, gt = be.t(g["x"]), be.t(g["bc"]), be.t(g.get("f")), be.t(g["gt"])
    xg = x.clone().requires_grad_(True)
    y = xg
    for _ in range(int(g["k"])):
        y = it(y, bc, f)
    loss = torch.nn.MSELoss()(y, gt)
    loss.backward()
    assert rel_err(be.n(y), g["y"]) <= 1e-5
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    if "w_first" in g:
        for ml, key in ((it.first_layers, "gw_first"), (it.pooling_layers, "gw_pool"), (it.second_layers, "gw_second")):
            gw = np.stack([be.n(l.weight.grad)[0, 0] for l in ml])
            assert rel_err(gw, g[key]) <= 5e-5, (key, gw, g[key])
    else:
        gw = np.stack([be.n(l.weight.grad)[0, 0] for l in it.layers])
        assert rel_err(gw, g["gw"]) <= 5e-5, (gw, g["gw"])
    assert rel_err(be.n(xg.grad), g["gx"]) <= 5e-5
    if "w_first" not in g:
        # the same k steps as ONE autograd node (npde_conv_iterate_tape / npde_conv_iterate_bwd): identical
        # forward, and gradients summed over the steps in the order autograd uses -> bit-identical
        chained = [be.n(l.weight.grad).copy() for l in it.layers], be.n(xg.grad).copy()
        for l in it.layers:
            l.weight.grad = None
        xg2 = x.clone().requires_grad_(True)
        y2 = it.unrolled(xg2, bc, f, int(g["k"]))
        torch.nn.MSELoss()(y2, gt).backward()
        assert_bitwise(be.n(y2), be.n(y), "fused unrolled forward")
        for l, ref in zip(it.layers, chained[0]):
            assert_bitwise(be.n(l.weight.grad), ref, "fused BPTT weight gradient")
        assert_bitwise(be.n(xg2.grad), chained[1], "fused BPTT input gradient")
        assert_bitwise(be.n(x), g["x"], "input untouched")


def test_full_bptt_training_option(be):
    """HeatModel(full_bptt=True) trains through every unrolled step; the default keeps the reference's
    last-step-only gradient.  Same start, same N: the two weight updates must differ, the losses must agree."""
    g = load_golden("driver_train_conv_sgd")
    outs = []
    for flag in (False, True):
        model = be.npde.HeatModel(make_opt(optimizer="sgd", lr_init=1e-2, max_iter_steps=5), full_bptt=flag)
        model.iterator.to(be.device)
        with torch.no_grad():
            model.iterator.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w0/")})
        np.random.seed(3)
        out = model.train(be.t(g["x"]), be.t(g["gt"]), be.t(g["bc"]), be.t(g["f"]))
        outs.append((out["loss"]["loss_x"], np.stack([be.n(l.weight)[0, 0] for l in model.iterator.layers])))
    assert abs(outs[0][0] - outs[1][0]) <= 1e-6 * abs(outs[0][0])
    assert np.abs(outs[0][1] - outs[1][1]).max() > 0


# ------------------------------------------------------------------- drivers --
def make_opt(**kw):
    import argparse
    d = dict(iterator="conv", activation="clamp", conv_n_layers=3, mg_n_layers=3, mg_pre_smoothing=2,
             mg_post_smoothing=2, cg_n_iters=4, geometry="square", is_train=True, optimizer="adam", lr_init=1e-3,
             max_iter_steps=6, max_iter_steps_from_gt=4, lambda_gt=0.0)
    d.update(kw)
    return argparse.Namespace(**d)


def test_get_iterator_contract(be):
    rows = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "get_iterator_contract.npy"),
                   allow_pickle=True)
    for itname, geom, name, cmp_name, ratio, is_train, is_mask, n_ops in rows:
        it, cmp_, r, tr = be.npde.get_iterator(make_opt(iterator=itname, geometry=geom))
        assert it.name() == name
        assert ("" if cmp_ is None else cmp_.name()) == cmp_name
        assert float(r) == float(ratio) and bool(tr) == bool(is_train)
        assert bool(it.is_bc_mask) == bool(is_mask) and float(it.n_operations) == float(n_ops)
    with pytest.raises(NotImplementedError):
        be.npde.get_iterator(make_opt(iterator="nope"))


def test_state_dict_keys(be):
    """Checkpoint compatibility (models/base_model.py:52-60): same keys and shapes as the reference modules."""
    c = be.npde.ConvIterator("clamp", 3)
    assert list(c.state_dict().keys()) == ["layers.0.weight", "layers.1.weight", "layers.2.weight"]
    assert all(tuple(v.shape) == (1, 1, 3, 3) for v in c.state_dict().values())
    u = be.npde.UNetIterator("clamp", 3, 2, 1)
    keys = list(u.state_dict().keys())
    assert keys == (["first_layers.%d.weight" % i for i in range(6)] + ["pooling_layers.%d.weight" % i for i in range(2)]
                    + ["second_layers.%d.weight" % i for i in range(3)])


def test_driver_evaluate(be):
    """HeatModel.evaluate / calculate_errors vs the reference run (golden): same error curves, same
    early-exit column, same final iterate."""
    g = load_golden("driver_evaluate_conv17")
    model = be.npde.HeatModel(make_opt(is_train=False))
    model.iterator.to(be.device)
    with torch.no_grad():
        for l, w in zip(model.iterator.layers, g["w"]):
            l.weight.copy_(torch.from_numpy(w).view(1, 1, 3, 3))
    results, xf = model.evaluate(be.t(g["x"]), be.t(g["gt"]), be.t(g["bc"]), None, 60)
    je, me = results["Jacobi errors"].numpy(), results["model errors"].numpy()
    assert je.shape == g["jacobi_errors"].shape and me.shape == g["model_errors"].shape
    # l2 reduction order differs from torch's mean: 1e-5 relative on the curves, exact zero-padding pattern
    assert np.array_equal(je == 0, g["jacobi_errors"] == 0) and np.array_equal(me == 0, g["model_errors"] == 0)
    assert rel_err(je, g["jacobi_errors"]) <= 1e-5 and rel_err(me, g["model_errors"]) <= 1e-5
    assert_bitwise(be.n(xf), g["x_final"], "final iterate of evaluate")


@pytest.mark.parametrize("tag", ["conv_adam", "conv_sgd", "unet_adam"])
def test_driver_train(be, tag):
    """HeatModel.train: three optimiser steps reproduce the reference's losses and updated weights."""
    g = load_golden("driver_train_" + tag)
    opt = {"conv_adam": make_opt(lambda_gt=0.3), "conv_sgd": make_opt(optimizer="sgd", lr_init=1e-2),
           "unet_adam": make_opt(iterator="unet", mg_n_layers=2, lambda_gt=0.0)}[tag]
    model = be.npde.HeatModel(opt)
    model.iterator.to(be.device)
    with torch.no_grad():
        model.iterator.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w0/")})
    np.random.seed(1234)
    x, gt, bc, f = be.t(g["x"]), be.t(g["gt"]), be.t(g["bc"]), be.t(g["f"])
    losses = []
    for _ in range(3):
        out = model.train(x.clone(), gt, bc, f)
        losses.append([out["loss"].get("loss_x", np.nan), out["loss"].get("loss_gt", np.nan)])
    losses = np.array(losses)
    ref = g["losses"]
    assert np.array_equal(np.isnan(losses), np.isnan(ref))
    assert np.allclose(losses[~np.isnan(ref)], ref[~np.isnan(ref)], rtol=2e-4)
    for k, v in g.items():
        if k.startswith("w1/"):
            got = model.iterator.state_dict()[k[3:]].cpu().numpy()
            assert np.allclose(got, v, rtol=0, atol=2e-5), (k, got, v)


def test_solve_to_tolerance_counts(be, oracle):
    """generation.py:72-99 as a device-side driver: identical iteration counts to the oracle's
    per-step loop (BASELINE: "identical iterations-to-tolerance"), Jacobi and Conv, tol 1e-4 at 17^2."""
    rng = np.random.RandomState(5)
    B, N = 5, 17
    bc = rng.rand(B, 4).astype(np.float32)
    x = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    w = np.zeros((2, 3, 3), np.float32)
    w[0, 1, 1], w[1, 1, 1] = 0.3, -0.3  # H = -0.09 * identity: Phi = 0.91 Psi + 0.09 I, a contraction
    tol = 1e-4
    for kind in ("jacobi", "conv"):
        if kind == "jacobi":
            it = be.npde.JacobiIterator()
        else:
            it = be.npde.ConvIterator("none", 2)
            with torch.no_grad():
                for l, wl in zip(it.layers, w):
                    l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
            it = it.to(be.device)
        res = be.utils.solve_to_tolerance(it, be.t(x), be.t(bc), None, tol=tol, check_every=50, max_iters=2000,
                                          per_problem=True)
        # oracle: step by step
        cur = x.copy()
        first = np.full(B, -1)
        n_done = 0
        for t in range(1, 2001):
            cur = oracle.fd_step(cur, bc, None) if kind == "jacobi" else oracle.conv_step(cur, bc, None, w, "none")
            e = oracle.fd_error(cur, bc, None, "max")
            for b in range(B):
                if first[b] < 0 and e[b] < tol:
                    first[b] = t
            if t % 50 == 0 and e.max() < tol:
                n_done = t
                break
        assert res["converged"] and res["iterations"] == n_done, (kind, res["iterations"], n_done)
        assert np.array_equal(res["first_below"].numpy(), first), (kind, res["first_below"], first)
        assert_bitwise(be.n(res["x"]), cur, "solution at the stopping iteration")
        # the cheap variant (no per-iteration rows) stops at the same check
        res2 = be.utils.solve_to_tolerance(it, be.t(x), be.t(bc), None, tol=tol, check_every=50, max_iters=2000)
        assert res2["iterations"] == n_done
        assert_bitwise(be.n(res2["x"]), cur, "solution, cheap variant")


# ------------------------------------------------- BASELINE.json full sizes (GPU) --
@pytest.mark.gpu
def test_full_size_headline_vs_oracle(oracle):
    """The bench workload itself (64 x 1025^2, Conv3, clamp): 3 fused steps and one single step against the
    CPU oracle, bit for bit, plus the residual row; then size-independent properties over 100 iterations."""
    be = Backend("cuda")
    rng = np.random.RandomState(666)
    B, N, L = 64, 1025, 3
    bc = rng.rand(B, 4).astype(np.float32)
    x = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    w = ((rng.rand(L, 3, 3) - 0.5) * 0.2).astype(np.float32)
    it = be.npde.ConvIterator("clamp", L)
    with torch.no_grad():
        for l, wl in zip(it.layers, w):
            l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
    it = it.to(be.device)
    xd, bcd = be.t(x), be.t(bc)
    with torch.no_grad():
        y3, _, efd = it.iterate(xd, bcd, None, 3, want_fd=True)
        y1 = it(xd, bcd, None)
    ref3, _, ref_fd = oracle.iterate("conv", x, bc, None, w, "clamp", 3, want_fd=True)
    assert_bitwise(be.n(y3), ref3, "3 fused Phi_H steps at 64 x 1025^2")
    assert_bitwise(be.n(efd), ref_fd, "fd_error rows at 64 x 1025^2")
    assert_bitwise(be.n(y1), oracle.conv_step(x, bc, None, w, "clamp"), "single step at 64 x 1025^2")
    with torch.no_grad():
        # k fused steps == k single steps (permuted fast path vs reference-layout path), 100 iterations
        y100, _, _ = it.iterate(xd, bcd, None, 100)
        z = xd
        for _ in range(100):
            z = it(z, bcd, None)
        assert torch.equal(y100, z)
        # problems are independent: a batch slice gives the slice of the batch result (what multi-GPU sharding relies on)
        ys, _, _ = it.iterate(xd[5:9].contiguous(), bcd[5:9].contiguous(), None, 100)
        assert torch.equal(ys, y100[5:9])
        # the Dirichlet ring is reproduced exactly, the clamp keeps the range
        assert torch.equal(y100[:, 0, 1:-1], bcd[:, 0:1].expand(-1, N - 2))
        assert torch.equal(y100[:, :, 0], bcd[:, 2:3].expand(-1, N))
        assert float(y100.min()) >= 0.0 and float(y100.max()) <= 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("B,N", [(2, 4097), (256, 257)])
def test_full_size_other_configs(oracle, B, N):
    """BASELINE configs 2 / 5 sizes: 4097^2 (square) and 256 x 257^2 (mask geometry): two fused steps of Conv3,
    four Jacobi steps (temporal depth 4 on the square domain) and one V-cycle against the oracle, bit for bit."""
    be = Backend("cuda")
    rng = np.random.RandomState(N)
    mask = N == 257
    if mask:
        bc = _mask_geometry(rng, B, N)
    else:
        bc = rng.rand(B, 4).astype(np.float32)
    x = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    w = ((rng.rand(3, 3, 3) - 0.5) * 0.2).astype(np.float32)
    it = be.npde.ConvIterator("clamp", 3)
    it.is_bc_mask = mask
    with torch.no_grad():
        for l, wl in zip(it.layers, w):
            l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
    it = it.to(be.device)
    jac = be.npde.JacobiIterator()
    jac.is_bc_mask = mask
    mg = be.npde.MultigridIterator(4, 2, 2)
    mg.is_bc_mask = mask
    xd, bcd = be.t(x), be.t(bc)
    with torch.no_grad():
        y2, _, _ = it.iterate(xd, bcd, None, 3)
        j4, _, _ = jac.iterate(xd, bcd, None, 4)
        v1 = mg(xd, bcd, None)
    assert_bitwise(be.n(y2), oracle.iterate("conv", x, bc, None, w, "clamp", 3)[0], "3 Conv3 steps")
    assert_bitwise(be.n(j4), oracle.iterate("jacobi", x, bc, None, None, "none", 4)[0], "4 Jacobi steps")
    assert_bitwise(be.n(v1), oracle.multigrid_step(x, bc, None, 4, 2, 2), "V-cycle (4,2,2)")


# --------------------------------------------------------------- error paths --
def test_error_behaviour(be):
    x = torch.zeros(2, 17, 17, device=be.device)
    with pytest.raises(Exception):  # heat_utils.py:40-42
        be.utils.fd_step(x, torch.zeros(2, 5, device=be.device), None)
    it = be.npde.ConvIterator("relu", 2).to(be.device)
    with pytest.raises(NotImplementedError):  # iterator.py:25-26
        it(x, torch.zeros(2, 4, device=be.device), None)
    with pytest.raises(NotImplementedError):  # heat_utils.py:130-131
        be.utils.fd_error(x, torch.zeros(2, 4, device=be.device), None, "sum")
    with pytest.raises(be.npde._lib.NpdeError):  # even multigrid level size
        be.npde.MultigridIterator(3, 1, 1)(torch.zeros(2, 19, 19, device=be.device),
                                           torch.zeros(2, 4, device=be.device), None)


def test_c_abi_argument_errors(be):
    """The C entry points reject bad arguments with the NPDE_ERR_* codes of include/npde.h instead of launching."""
    import ctypes
    L = be.npde._lib
    lib = L.lib()
    x = torch.zeros(2, 17, 17, device=be.device)
    y = torch.zeros_like(x)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    assert lib.npde_cg_step(p(x), 3, p(x), 2, 17, None, 0, None) == -7          # in place: NPDE_ERR_ARG
    assert lib.npde_cg_step(p(x), 3, p(y), 2, 17, None, 0, None) == -6          # no workspace: NPDE_ERR_WORKSPACE
    assert lib.npde_cg_step(None, 3, p(y), 2, 17, None, 0, None) == -1          # NPDE_ERR_NULL
    assert lib.npde_cg_step(p(x), 3, p(y), 0, 17, None, 0, None) == -2          # NPDE_ERR_SIZE
    bc = torch.zeros(2, 4, device=be.device)
    w = torch.zeros(3, 3, 3, device=be.device)
    assert lib.npde_conv_iterate_tape(p(x), p(bc), 0, None, p(w), None, 3, 1, 0, p(y), p(y), 2, 17, None, 0, None) == -7
    assert lib.npde_conv_iterate_tape(p(x), p(bc), 0, None, p(w), None, 3, 1, 2, None, p(y), 2, 17, None, 0, None) == -1
    assert lib.npde_conv_iterate_bwd(p(x), p(bc), 0, None, p(w), 3, 7, 2, p(y), p(w), None, 2, 17, None, 0, None) == -4
    assert lib.npde_conv_iterate_bwd(p(x), p(bc), 0, None, p(w), 9, 1, 2, p(y), p(w), None, 2, 17, None, 0, None) == -5
    assert lib.npde_conv_iterate_bwd(p(x), p(bc), 5, None, p(w), 3, 1, 2, p(y), p(w), None, 2, 17, None, 0, None) == -3
    assert lib.npde_strerror(-6).decode().startswith("workspace")
    assert lib.npde_workspace_bytes(L.WS_CG, 2, 17, 0, 0, 0, 0, 0) > 4 * 2 * 17 * 17 * 4
    assert lib.npde_workspace_bytes(L.WS_CONV_BPTT, 2, 17, 0, 3, 0, 0, 0) > \
        lib.npde_workspace_bytes(L.WS_CONV_BWD, 2, 17, 0, 3, 0, 0, 0)


@pytest.mark.parametrize("B,N,k", [(3, 65, 9), (2, 129, 12), (5, 33, 7)])
def test_mask_jacobi_temporal_blocking_open_border(be, oracle, B, N, k):
    """Temporally blocked Jacobi on a mask geometry whose outer ring is NOT Dirichlet (m = 0 there): the
    zero padding beyond the grid then matters at every stage.  Depths 1, 2, 4 and the default vs the oracle."""
    rng = np.random.RandomState(N + k)
    m = (rng.rand(B, N, N) < 0.08).astype(np.float32)
    v = rng.rand(B, N, N).astype(np.float32) * m
    bc = np.stack([v, m], 1)
    x = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    ref, _, ref_fd = oracle.iterate("jacobi", x, bc, None, None, "none", k, want_fd=True)
    jac = be.npde.JacobiIterator()
    jac.is_bc_mask = True
    xd, bcd = be.t(x), be.t(bc)
    for depth in (0, 1, 2, 4):
        y, _, _ = jac.iterate(xd, bcd, None, k, temporal_depth=depth)
        assert_bitwise(be.n(y), ref, "mask Jacobi, %d steps, depth %d" % (k, depth))
    y, _, fd = jac.iterate(xd, bcd, None, k, want_fd=True)
    assert_bitwise(be.n(y), ref, "with residual rows")
    assert_bitwise(be.n(fd), ref_fd, "residual rows")
    # the fused Conv kernel and the V-cycle on the same open-border geometry
    w = ((rng.rand(3, 3, 3) - 0.5) * 0.3).astype(np.float32)
    it = be.npde.ConvIterator("clamp", 3)
    it.is_bc_mask = True
    with torch.no_grad():
        for l, wl in zip(it.layers, w):
            l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
    it = it.to(be.device)
    yc, _, _ = it.iterate(xd, bcd, None, 5)
    assert_bitwise(be.n(yc), oracle.iterate("conv", x, bc, None, w, "clamp", 5)[0], "5 Conv3 steps, open border")
    if N >= 65:
        mg = be.npde.MultigridIterator(3, 3, 5)
        mg.is_bc_mask = True
        assert_bitwise(be.n(mg(xd, bcd, None)), oracle.multigrid_step(x, bc, None, 3, 3, 5), "V-cycle (3,3,5), open border")


@pytest.mark.parametrize("B,N,L,mask,has_f,act,k", [(3, 65, 3, False, False, "clamp", 7), (2, 33, 2, True, True, "none", 5),
                                                    (5, 17, 1, False, True, "sigmoid", 4), (2, 65, 4, True, False, "clamp", 6),
                                                    (1, 5, 3, False, False, "clamp", 3)])
def test_small_grid_on_chip_loop(be, oracle, B, N, L, mask, has_f, act, k):
    """k_small_iterate (grids up to 65^2: the whole loop in one kernel, one CTA per problem): iterates and both
    kinds of error rows against the oracle, Conv and Jacobi, square and mask geometries, with and without source."""
    rng = np.random.RandomState(B * 131 + N + L)
    bc = _mask_geometry(rng, B, N) if mask else rng.rand(B, 4).astype(np.float32)
    x = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    gt = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    f = None
    if has_f:
        f = np.zeros((B, N, N), np.float32)
        f[:, 1:-1, 1:-1] = (rng.rand(B, N - 2, N - 2) - 0.5) * 1e-3
    w = ((rng.rand(L, 3, 3) - 0.5) * 0.3).astype(np.float32)
    it = be.npde.ConvIterator(act, L)
    it.is_bc_mask = mask
    with torch.no_grad():
        for l, wl in zip(it.layers, w):
            l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
    it = it.to(be.device)
    xd, bcd, fd, gtd = be.t(x), be.t(bc), be.t(f), be.t(gt)
    y, e2, ef = it.iterate(xd, bcd, fd, k, gt=gtd, want_l2=True, want_fd=True)
    ref, r2, rf = oracle.iterate("conv", x, bc, f, w, act, k, gt=gt, want_fd=True)
    if act == "sigmoid":
        assert rel_err(be.n(y), ref) <= 1e-5
    else:
        assert_bitwise(be.n(y), ref, "on-chip Conv loop")
        assert_bitwise(be.n(ef), rf, "fd_error rows")
        assert rel_err(be.n(e2), r2) <= 1e-6
    jac = be.npde.JacobiIterator()
    jac.is_bc_mask = mask
    yj, j2, jf = jac.iterate(xd, bcd, fd, k + 3, gt=gtd, want_l2=True, want_fd=True)
    refj, rj2, rjf = oracle.iterate("jacobi", x, bc, f, None, "none", k + 3, gt=gt, want_fd=True)
    assert_bitwise(be.n(yj), refj, "on-chip Jacobi loop")
    assert_bitwise(be.n(jf), rjf, "Jacobi fd_error rows")
    assert rel_err(be.n(j2), rj2) <= 1e-6
    assert_bitwise(be.n(xd), x, "input untouched")
    # in place
    xc = xd.clone()
    jac.iterate(xc, bcd, fd, k + 3, out=xc)
    assert_bitwise(be.n(xc), refj, "in place")


@pytest.mark.gpu
def test_large_batch_layered_kernels_vs_oracle(oracle):
    """Sizes at which the coarse U-Net levels and the backward run on the row-sliding kernels (k_conv3x3_same,
    k_bwd_layer; >= 2^20 cells per launch): U-Net step on a mask geometry bit for bit, Conv3 backward to 1e-5."""
    be = Backend("cuda")
    rng = np.random.RandomState(77)
    B, N = 20, 513
    bc = _mask_geometry(rng, B, N)
    x = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    wf = ((rng.rand(6, 3, 3) - 0.5) * 0.3).astype(np.float32)
    wp = ((rng.rand(2, 3, 3) - 0.5) * 0.3).astype(np.float32)
    ws = ((rng.rand(3, 3, 3) - 0.5) * 0.3).astype(np.float32)
    it = be.npde.UNetIterator("clamp", 3, 2, 1)
    it.is_bc_mask = True
    with torch.no_grad():
        for ml, wsrc in ((it.first_layers, wf), (it.pooling_layers, wp), (it.second_layers, ws)):
            for l, wl in zip(ml, wsrc):
                l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
    it = it.to(be.device)
    with torch.no_grad():
        y = it(be.t(x), be.t(bc), None)
    assert_bitwise(be.n(y), oracle.unet_step(x, bc, None, wf, wp, ws, 3, 2, 1, "clamp"), "U-Net step, 20 x 513^2 masks")
    # Conv3 one-step backward on the square domain
    bc4 = rng.rand(B, 4).astype(np.float32)
    xs = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc4)
    w = ((rng.rand(3, 3, 3) - 0.5) * 0.3).astype(np.float32)
    gout = (rng.rand(B, N, N).astype(np.float32) - 0.5)
    gw, gx = be.F.conv_step_bwd(be.t(xs), be.t(bc4), None, be.t(w), "clamp", be.t(gout), True)
    rw, rx = oracle.conv_step_bwd(xs, bc4, None, w, "clamp", gout)
    assert rel_err(be.n(gw), rw) <= 1e-5
    assert rel_err(be.n(gx), rx) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("mask,has_f", [(False, False), (True, True), (False, True), (True, False)])
def test_graph_replay_paths_vs_oracle(oracle, mask, has_f):
    """npde_iterate on a launch-bound size that takes the CUDA-graph replay (129^2, k >= 16): captured segments of 32
    fused launches with and without staged error rows, Conv and temporally blocked Jacobi, U-Net / multigrid pairs of
    steps -- all against the oracle."""
    be = Backend("cuda")
    rng = np.random.RandomState(5 + 2 * mask + has_f)
    B, N, k = 3, 129, 75
    bc = _mask_geometry(rng, B, N) if mask else rng.rand(B, 4).astype(np.float32)
    x = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    gt = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    f = None
    if has_f:
        f = np.zeros((B, N, N), np.float32)
        f[:, 1:-1, 1:-1] = (rng.rand(B, N - 2, N - 2) - 0.5) * 1e-3
    w = ((rng.rand(3, 3, 3) - 0.5) * 0.2).astype(np.float32)
    it = be.npde.ConvIterator("clamp", 3)
    it.is_bc_mask = mask
    with torch.no_grad():
        for l, wl in zip(it.layers, w):
            l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
    it = it.to(be.device)
    xd, bcd, fd, gtd = be.t(x), be.t(bc), be.t(f), be.t(gt)
    ref, r2, rf = oracle.iterate("conv", x, bc, f, w, "clamp", k, gt=gt, want_fd=True)
    y, _, _ = it.iterate(xd, bcd, fd, k)
    assert_bitwise(be.n(y), ref, "Conv3, %d steps, replayed segments" % k)
    y, e2, ef = it.iterate(xd, bcd, fd, k, gt=gtd, want_l2=True, want_fd=True)
    assert_bitwise(be.n(y), ref, "Conv3 with staged error rows")
    assert_bitwise(be.n(ef), rf, "fd_error rows")
    assert rel_err(be.n(e2), r2) <= 1e-6
    jac = be.npde.JacobiIterator()
    jac.is_bc_mask = mask
    kj = 2 * k + 1
    refj, _, rjf = oracle.iterate("jacobi", x, bc, f, None, "none", kj, want_fd=True)
    yj, _, _ = jac.iterate(xd, bcd, fd, kj)
    assert_bitwise(be.n(yj), refj, "Jacobi, %d steps, replayed segments of blocked launches" % kj)
    yj, _, jf = jac.iterate(xd, bcd, fd, kj, want_fd=True)
    assert_bitwise(be.n(yj), refj, "Jacobi with staged residual rows")
    assert_bitwise(be.n(jf), rjf, "Jacobi fd_error rows")
    mg = be.npde.MultigridIterator(3, 2, 2)
    mg.is_bc_mask = mask
    ym, _, _ = mg.iterate(xd, bcd, fd, 17)
    refm = x
    for _ in range(17):
        refm = oracle.multigrid_step(refm, bc, f, 3, 2, 2)
    assert_bitwise(be.n(ym), refm, "17 V-cycles, replayed pairs")
    wf = ((rng.rand(4, 3, 3) - 0.5) * 0.2).astype(np.float32)
    wp = ((rng.rand(1, 3, 3) - 0.5) * 0.2).astype(np.float32)
    ws = ((rng.rand(2, 3, 3) - 0.5) * 0.2).astype(np.float32)
    un = be.npde.UNetIterator("clamp", 2, 2, 1)
    un.is_bc_mask = mask
    with torch.no_grad():
        for ml, wsrc in ((un.first_layers, wf), (un.pooling_layers, wp), (un.second_layers, ws)):
            for l, wl in zip(ml, wsrc):
                l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
    un = un.to(be.device)
    yu, _, _ = un.iterate(xd, bcd, fd, 18)
    refu = x
    for _ in range(18):
        refu = oracle.unet_step(refu, bc, f, wf, wp, ws, 2, 2, 1, "clamp")
    assert_bitwise(be.n(yu), refu, "18 U-Net steps, replayed pairs")


# ------------------------------------------ second-generation Phi_H kernel (npde_wide.cu) --
@pytest.mark.parametrize("B,N,L,has_f,act,k", [(3, 129, 3, False, "clamp", 5), (2, 257, 3, True, "none", 4),
                                               (7, 33, 2, True, "clamp", 6), (13, 17, 1, False, "clamp", 3),
                                               (4, 9, 4, False, "none", 3), (1, 65, 3, False, "sigmoid", 3),
                                               (31, 17, 3, True, "clamp", 4), (2, 12, 3, False, "clamp", 3),
                                               (1, 513, 2, False, "clamp", 3)])
def test_wide_kernel_vs_oracle(be, oracle, B, N, L, has_f, act, k, tmp_path, monkeypatch):
    """k_conv_wide: batch-interleaved W8 layout, eight columns per lane, strips that span several images and
    pieces that straddle strips (the emulator build has 8 warp slots: several bands per strip; 1 x 513^2 has an
    interior strip between two edge strips, which get one band more than it).  k fused steps
    through npde_iterate (pack -> k launches -> unpack) against the CPU oracle, bit for bit (sigmoid: 1e-5)."""
    rng = np.random.RandomState(B * 1000 + N + L)
    bc = rng.rand(B, 4).astype(np.float32)
    x = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    f = None
    if has_f:
        f = np.zeros((B, N, N), np.float32)
        f[:, 1:-1, 1:-1] = (rng.rand(B, N - 2, N - 2) - 0.5) * 1e-3
    w = ((rng.rand(L, 3, 3) - 0.5) * 0.3).astype(np.float32)
    it = be.npde.ConvIterator(act, L)
    with torch.no_grad():
        for l, wl in zip(it.layers, w):
            l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
    it = it.to(be.device)
    xd, bcd, fd = be.t(x), be.t(bc), be.t(f)
    trace = tmp_path / "launches.txt"
    if be.kind == "emu":
        monkeypatch.setenv("NPDE_EMU_TRACE", str(trace))
    else:
        monkeypatch.setenv("NPDE_WIDE", "1")
    with torch.no_grad():
        yk = be.F.iterate(be.npde._lib.IT_CONV, xd, bcd, fd, k, w=torch.stack([l.weight.reshape(9) for l in it.layers]),
                          w_host=w, n_layers=L, act=act, temporal_depth=-1)[0]
    ref = oracle.iterate("conv", x, bc, f, w, act, k)[0]
    if act == "sigmoid":
        assert rel_err(be.n(yk), ref) <= 1e-5
    else:
        assert_bitwise(be.n(yk), ref, "%d wide Phi_H steps" % k)
    if be.kind == "emu":
        names = trace.read_text().split()
        assert sum("k_conv_wide" in n for n in names) == k, names
        assert sum("k_wide_convert" in n for n in names) == (3 if has_f else 2), names


@pytest.mark.parametrize("B,N,L,has_f,k", [(3, 65, 3, False, 5), (2, 33, 2, True, 4), (5, 17, 1, False, 3)])
def test_packed_solver_loop(be, oracle, B, N, L, has_f, k):
    """The persistent packed layout (npde_pack / npde_conv_iterate_packed / npde_fd_error_packed / npde_unpack):
    pack once, k steps, residual on the packed field, unpack -- iterate and residual bit-identical to the oracle's
    and to the drop-in path on reference-layout tensors."""
    rng = np.random.RandomState(77 + B + N)
    bc = rng.rand(B, 4).astype(np.float32)
    x = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    f = None
    if has_f:
        f = np.zeros((B, N, N), np.float32)
        f[:, 1:-1, 1:-1] = (rng.rand(B, N - 2, N - 2) - 0.5) * 1e-3
    w = ((rng.rand(L, 3, 3) - 0.5) * 0.3).astype(np.float32)
    it = be.npde.ConvIterator("clamp", L)
    with torch.no_grad():
        for l, wl in zip(it.layers, w):
            l.weight.copy_(torch.from_numpy(wl).view(1, 1, 3, 3))
    it = it.to(be.device)
    xd, bcd, fd = be.t(x), be.t(bc), be.t(f)
    with torch.no_grad():
        a = it.pack(xd)
        fp = it.pack(fd) if has_f else None
        b = a.empty_like()
        assert_bitwise(be.n(it.unpack(a)), x, "pack -> unpack")
        res = it.iterate_packed(a, b, bcd, fp, k)
        assert res is (a if k % 2 == 0 else b)
        err = it.fd_error_packed(res, fp)
        y = it.unpack(res)
    ref = oracle.iterate("conv", x, bc, f, w, "clamp", k)[0]
    assert_bitwise(be.n(y), ref, "%d packed Phi_H steps" % k)
    assert_bitwise(be.n(err), oracle.fd_error(ref, bc, f, "max"), "fd_error on the packed field")
    assert_bitwise(be.n(err), be.n(be.utils.fd_error(y, bcd, fd)), "packed vs unpacked fd_error")


# ----------------------------------------------------------------- initialize --
@pytest.mark.parametrize("name", golden_names("init_"))
def test_initialize_vs_reference(be, name):
    """utils.initialize (utils/heat_utils.py:70-94) through npde_initialize against the reference's own outputs:
    all modes, both bc encodings, bit for bit -- 'random' included (same torch generator, same draw)."""
    g = load_golden(name)
    bc = be.t(g["bc"])
    for mode in ("zero", "random", "avg"):
        x = be.t(g["x"]).clone()
        torch.manual_seed(int(g["seed"][0]))
        y = be.utils.initialize(x, bc, mode)
        assert_bitwise(be.n(y), g[mode], "initialize(%s)" % mode)
        if g["bc"].ndim == 2:  # square domain: in place (heat_utils.py:85-91)
            assert y.data_ptr() == x.data_ptr()
    if g["bc"].ndim == 2:
        with pytest.raises(NotImplementedError):
            be.utils.initialize(be.t(g["x"]).clone(), bc, "bogus")


def test_initialize_on_host_batches():
    """The reference's drivers initialise the CPU batch of the DataLoader before the model moves it to the GPU
    (train.py:29, test.py:31): host tensors take the reference's torch expressions, every mode, both encodings."""
    npde = emu_backend.use_cuda_library()
    for name in golden_names("init_"):
        g = load_golden(name)
        bc = torch.from_numpy(g["bc"])
        for mode in ("zero", "random", "avg"):
            torch.manual_seed(int(g["seed"][0]))
            y = npde.utils.initialize(torch.from_numpy(g["x"]).clone(), bc, mode)
            assert not y.is_cuda
            assert_bitwise(y.numpy(), g[mode], "host initialize(%s) %s" % (mode, name))


# ------------------------ iterators trained by the reference: BASELINE configs[0] + iterations to 1e-6 --
def _load_trained(be, tag):
    g = load_golden("trained_" + tag)
    L, levels, pre, post = (int(v) for v in g["meta"])
    opt = make_opt(is_train=False) if tag == "conv3" else \
        make_opt(iterator="unet", mg_n_layers=levels, mg_pre_smoothing=pre, mg_post_smoothing=post, is_train=False)
    model = be.npde.HeatModel(opt)
    model.iterator.to(be.device)
    with torch.no_grad():
        model.iterator.load_state_dict({k[2:].replace("/", "."): torch.from_numpy(v)
                                        for k, v in g.items() if k.startswith("w/")})
    return g, model


@pytest.mark.parametrize("tag", ["conv3", "unet322"])
def test_config1_evaluate_with_trained_iterator(be, tag):
    """BASELINE configs[0]: 32 x 65^2, 800 evaluation steps from the 'avg' start through HeatModel.evaluate
    (test.py:93-120, heat_model.py:92-128) with weights TRAINED BY THE REFERENCE; the reference's own error
    curves, final iterate and Metrics.  The emulator runs the first 3 problems for 40 steps (problems are
    independent and the kernels batch-invariant), the GPU the whole configuration."""
    g, model = _load_trained(be, tag)
    nb, steps = (32, 800) if be.kind == "cuda" else (3, 40)
    x, gt, bc = be.t(g["x"][:nb]), be.t(g["gt"][:nb]), be.t(g["bc"][:nb])
    results, xf = model.evaluate(x.clone(), gt, bc, None, steps)
    me, je = be.n(results["model errors"]), be.n(results["Jacobi errors"])
    ref_m, ref_j = g["model_errors"][:nb, :steps + 1], g["jacobi_errors"][:nb, :steps + 1]
    # the reference stops filling a row block once every problem is below the threshold (heat_utils.py:173-176):
    # the same columns are zero; non-zero entries agree to the fp32 tolerance of the mean reduction
    assert np.array_equal(me == 0, ref_m == 0) or nb < 32
    live = (ref_m != 0) & (me != 0)
    assert np.abs(me[live] - ref_m[live]).max() <= 1e-5 * max(np.abs(ref_m[live]).max(), 1e-30)
    livej = (ref_j != 0) & (je != 0)
    assert np.abs(je[livej] - ref_j[livej]).max() <= 1e-5 * max(np.abs(ref_j[livej]).max(), 1e-30)
    if nb == 32:
        assert_bitwise(be.n(xf), g["x_final"], "final iterate of the 800-step evaluation") if tag == "conv3" else None
        m = be.npde.utils.Metrics(scale=1, error_threshold=0.01)
        m.update(results)
        got = m.get_results()
        ref = g["metrics"]
        for key, r in zip(("Jacobi", "model", "ratio"), ref):
            assert (np.isnan(r) and np.isnan(got[key])) or np.isclose(got[key], r, rtol=1e-6), (key, got[key], r)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["conv3", "unet322"])
def test_iterations_to_1e6_match_the_reference(tag):
    """North star: "bit-for-bit identical iteration counts to reach a 1e-6 residual".  Per-problem first iteration
    with fd_error < 1e-6 at 65^2, for the iterator trained by the reference and for Jacobi, against the counts the
    REFERENCE's own iterators needed (tests/golden/trained_*.npz, 8 problems from the 'avg' start)."""
    be = Backend("cuda")
    g, model = _load_trained(be, tag)
    x, bc = be.t(g["x"][:8]), be.t(g["bc"][:8])
    with torch.no_grad():
        res = be.utils.solve_to_tolerance(model.iterator, x, bc, None, tol=1e-6, check_every=50, max_iters=5000,
                                          per_problem=True)
        resj = be.utils.solve_to_tolerance(be.npde.JacobiIterator(), x, bc, None, tol=1e-6, check_every=500,
                                           max_iters=40000, per_problem=True)
    assert res["converged"] and resj["converged"]
    assert np.array_equal(res["first_below"].numpy(), g["first_model"]), (res["first_below"], g["first_model"])
    assert np.array_equal(resj["first_below"].numpy(), g["first_jacobi"]), (resj["first_below"], g["first_jacobi"])


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,kind", [(4, 129, "jacobi"), (4, 129, "conv3"), (2, 257, "jacobi"), (3, 257, "conv3")])
def test_iterations_to_1e6_match_the_oracle(oracle, B, N, kind):
    """Same identity at 129^2 / 257^2 against the CPU oracle (the reference itself would take minutes): Jacobi and
    the reference-trained Conv3 from the 'avg' start to fd_error < 1e-6, per-problem first crossing and the
    iterate at the stopping check."""
    be = Backend("cuda")
    rng = np.random.RandomState(N + B)
    bc = rng.rand(B, 4).astype(np.float32)
    x = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    x[:, 1:-1, 1:-1] = (((bc[:, 0] + bc[:, 1]) + bc[:, 2]) + bc[:, 3])[:, None, None] * np.float32(0.25)  # 'avg'
    g, model = _load_trained(be, "conv3")
    w = np.stack([g["w/layers/%d/weight" % i][0, 0] for i in range(3)])
    it = be.npde.JacobiIterator() if kind == "jacobi" else model.iterator
    every, cap = (1000, 200000) if kind == "jacobi" else (100, 20000)
    with torch.no_grad():
        res = be.utils.solve_to_tolerance(it, be.t(x), be.t(bc), None, tol=1e-6, check_every=every, max_iters=cap,
                                          per_problem=True)
    assert res["converged"]
    n = res["iterations"]
    ref, _, rows = oracle.iterate("jacobi" if kind == "jacobi" else "conv", x, bc, None,
                                  None if kind == "jacobi" else w, "none" if kind == "jacobi" else "clamp", n,
                                  want_fd=True)
    below = rows[1:] < 1e-6
    first = np.array([int(np.argmax(below[:, b])) + 1 if below[:, b].any() else -1 for b in range(B)])
    assert np.array_equal(res["first_below"].numpy(), first), (res["first_below"], first)
    assert rows[n].max() < 1e-6 and (n == every or rows[n - every].max() >= 1e-6), "stops at the same check"
    assert_bitwise(be.n(res["x"]), ref, "iterate at the stopping check")


# ------------------------------------------------------- device-side solve loop (npde_solve) --
@pytest.mark.parametrize("case", ["jacobi_square_f", "jacobi_mask", "conv_mask", "multigrid_cap", "unet_square",
                                  "already_converged"])
def test_device_side_solve_loop(be, oracle, case):
    """npde_solve (one CUDA graph with a WHILE node; the residual kernel writes the loop condition on the device)
    against the host-driven loop of solve_to_tolerance: same stopping check, same iteration count, same iterate
    bit for bit, same residuals -- Jacobi / Conv / Multigrid / U-Net, square domain and mask geometry, with a source
    term, a run that hits max_iters, and an input that already meets the tolerance."""
    rng = np.random.RandomState(11)
    B, N = 3, 33
    npde = be.npde
    f = None
    tol, every, cap = 1e-4, 20, 4000
    if case in ("jacobi_mask", "conv_mask"):
        bc = _mask_geometry(rng, B, N)
        x = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    else:
        bc = rng.rand(B, 4).astype(np.float32)
        x = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    if case == "jacobi_square_f":
        f = np.zeros((B, N, N), np.float32)
        f[:, 1:-1, 1:-1] = (rng.rand(B, N - 2, N - 2) - 0.5) * 1e-3
    if case.startswith("jacobi") or case == "already_converged":
        it = npde.JacobiIterator()
    elif case == "conv_mask":
        it = npde.ConvIterator("clamp", 2)
        with torch.no_grad():
            it.layers[0].weight.zero_(); it.layers[1].weight.zero_()
            it.layers[0].weight[0, 0, 1, 1] = 0.3; it.layers[1].weight[0, 0, 1, 1] = -0.3
        it = it.to(be.device)
    elif case == "multigrid_cap":
        it, tol, every, cap = npde.MultigridIterator(3, 2, 2), 1e-7, 4, 40  # the interpolant keeps it above 1e-7
    else:
        g, model = _load_trained(be, "unet322")
        it, tol, every, cap = model.iterator, 1e-5, 10, 2000
    it.is_bc_mask = bc.ndim == 4
    xd, bcd, fd = be.t(x), be.t(bc), be.t(f)
    if case == "already_converged":
        xd = be.utils.solve_to_tolerance(it, xd, bcd, fd, tol=tol, check_every=every, max_iters=cap,
                                         device_loop=False)["x"]
    with torch.no_grad():
        host = be.utils.solve_to_tolerance(it, xd, bcd, fd, tol=tol, check_every=every, max_iters=cap, device_loop=False)
        dev = be.utils.solve_to_tolerance(it, xd, bcd, fd, tol=tol, check_every=every, max_iters=cap)
    assert dev["iterations"] == host["iterations"], (case, dev["iterations"], host["iterations"])
    assert dev["converged"] == host["converged"] == (case != "multigrid_cap")
    assert_bitwise(be.n(dev["x"]), be.n(host["x"]), "iterate at the stopping check")
    assert dev["history"][0][1] == host["history"][0][1] and dev["history"][-1][1] == host["history"][-1][1]
    if case == "already_converged":
        assert dev["iterations"] == 0
    if case == "multigrid_cap":
        assert dev["iterations"] == cap
