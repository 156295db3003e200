"""Pin the CPU oracle (oracle/npde_oracle.c) against golden vectors produced by
the unmodified reference (tests/golden/make_golden.py).

Bar: bit-exact wherever the reference's ATen summation order is reproducible
(every case with batch >= 2, except sigmoid and the mean reductions), else the
north-star tolerance 1e-5 relative, written next to each assert.
"""
import numpy as np
import pytest

from conftest import assert_bitwise, golden_names, load_golden, rel_err

ACTS = {0: "none", 1: "clamp", 2: "sigmoid"}


@pytest.mark.parametrize("name", golden_names("ops_"))
def test_grid_operators(oracle, name):
    g = load_golden(name)
    x, bc, gt = g["x"], g["bc"], g["gt"]
    assert_bitwise(oracle.fd_step(x, bc), g["fd_step"], "fd_step")
    assert_bitwise(oracle.fd_error(x, bc, None, "max"), g["fd_error_max"], "fd_error max")
    assert rel_err(oracle.fd_error(x, bc, None, "mean"), g["fd_error_mean"]) < 1e-6
    if "f" in g:
        assert_bitwise(oracle.fd_step(x, bc, g["f"]), g["fd_step_f"], "fd_step f")
        assert_bitwise(oracle.fd_error(x, bc, g["f"], "max"), g["fd_error_max_f"], "fd_error f")
        assert rel_err(oracle.fd_error(x, bc, g["f"], "mean"), g["fd_error_mean_f"]) < 1e-6
    assert rel_err(oracle.l2_error(x, gt), g["l2_error"]) < 1e-6
    assert_bitwise(oracle.set_boundary(gt, bc), g["set_boundary"], "set_boundary")
    assert_bitwise(oracle.pad_boundary(gt[:, 1:-1, 1:-1], bc), g["pad_boundary"], "pad_boundary")
    y, _, ef = oracle.iterate("jacobi", x, bc, g.get("f"), None, "none", 25, want_fd=True)
    assert_bitwise(y, g["jacobi25"], "25 chained Jacobi steps")
    assert_bitwise(ef[-1], g["jacobi25_fd_error"], "fd_error after 25 steps")
    assert_bitwise(oracle.subsample(x), g["subsample"], "subsample")
    if bc.ndim == 4:
        B, _, N, _ = bc.shape
        assert_bitwise(oracle.subsample(bc.reshape(2 * B, N, N)).reshape(g["bc_sub"].shape),
                       g["bc_sub"], "bc subsample")
    assert_bitwise(oracle.restriction(x, g["bc_sub"]), g["restriction"], "restriction")
    assert_bitwise(oracle.interpolation(g["xs"], bc), g["interpolation"], "interpolation")


@pytest.mark.parametrize("name", golden_names("conv_"))
def test_conv_iterator(oracle, name):
    g = load_golden(name)
    L, act, is_mask = (int(v) for v in g["meta"])
    x, bc, f, w = g["x"], g["bc"], g.get("f"), g["w"]
    y1 = oracle.conv_step(x, bc, f, w, ACTS[act])
    h = oracle.conv_H(g["r"], bc, w, bool(is_mask))
    y10, _, _ = oracle.iterate("conv", x, bc, f, w, ACTS[act], 10)
    if x.shape[0] >= 2 and ACTS[act] != "sigmoid":
        assert_bitwise(h, g["H"], "H")
        assert_bitwise(y1, g["step1"], "conv step")
        assert_bitwise(y10, g["step10"], "10 conv steps")
    else:  # B == 1 changes ATen's conv order; sigmoid uses a vectorised exp
        assert rel_err(h, g["H"]) <= 1e-5
        assert rel_err(y1, g["step1"]) <= 1e-5
        assert rel_err(y10, g["step10"]) <= 1e-5


@pytest.mark.parametrize("name", golden_names("mg_"))
def test_multigrid_iterator(oracle, name):
    g = load_golden(name)
    nl, pre, post, _ = (int(v) for v in g["meta"])
    y1 = oracle.multigrid_step(g["x"], g["bc"], g.get("f"), nl, pre, post)
    assert_bitwise(y1, g["step1"], "V-cycle")
    y3 = oracle.multigrid_step(oracle.multigrid_step(y1, g["bc"], g.get("f"), nl, pre, post),
                               g["bc"], g.get("f"), nl, pre, post)
    assert_bitwise(y3, g["step3"], "3 V-cycles")


@pytest.mark.parametrize("name", golden_names("unet_"))
def test_unet_iterator(oracle, name):
    g = load_golden(name)
    nl, pre, post, act, is_mask = (int(v) for v in g["meta"])
    args = (g["w_first"], g["w_pool"], g["w_second"], nl, pre, post)
    h = oracle.unet_H(g["r"], g["bc"], *args, bool(is_mask))
    y1 = oracle.unet_step(g["x"], g["bc"], g.get("f"), *args, ACTS[act])
    y = g["x"]
    for _ in range(5):
        y = oracle.unet_step(y, g["bc"], g.get("f"), *args, ACTS[act])
    assert_bitwise(h, g["H"], "unet H")
    assert_bitwise(y1, g["step1"], "unet step")
    assert_bitwise(y, g["step5"], "5 unet steps")


@pytest.mark.parametrize("name", golden_names("bwd_conv_"))
def test_conv_backward(oracle, name):
    g = load_golden(name)
    L, act, _ = (int(v) for v in g["meta"])
    x, bc, f, w, gt = g["x"], g["bc"], g.get("f"), g["w"], g["gt"]
    y = oracle.conv_step(x, bc, f, w, ACTS[act])
    assert rel_err(y, g["y"]) <= 1e-5
    gout = (2.0 * (y.astype(np.float64) - gt) / y.size).astype(np.float32)  # d MSE / d y
    gw, gx = oracle.conv_step_bwd(x, bc, f, w, ACTS[act], gout)
    # autograd's reduction order is unspecified: tolerance, relative to the largest entry
    assert rel_err(gw, g["gw"]) <= 2e-5, (gw, g["gw"])
    assert rel_err(gx, g["gx"]) <= 2e-5


@pytest.mark.parametrize("name", golden_names("cg_"))
def test_conjugate_gradient(oracle, name):
    """ConjugateGradient.forward: the dot products are torch.sum trees of unspecified order, so the oracle
    (fp64-accumulated sums) matches the reference to rounding: 1e-5 relative."""
    g = load_golden(name)
    n_iters = int(g["meta"][0])
    y = oracle.cg_step(g["x"], n_iters)
    assert rel_err(y, g["y"]) <= 1e-5
    assert rel_err(oracle.cg_step(y, n_iters), g["y2"]) <= 1e-5
    assert_bitwise(y[:, 0, :], g["x"][:, 0, :], "boundary ring kept")


@pytest.mark.parametrize("name", golden_names("mgres_"))
def test_multigrid_residual(oracle, name):
    g = load_golden(name)
    L, pre, post = (int(v) for v in g["meta"])
    assert_bitwise(oracle.multigrid_residual_step(g["x"], g["bc"], g.get("f"), L, pre, post), g["y"],
                   "MultigridResidualIterator.forward")
