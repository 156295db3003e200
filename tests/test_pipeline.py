"""The callers and data formats either side of the hot path: problem generators, the
solve-to-tolerance dataset pipeline and its on-disk format, the dataset readers, the runtime.py protocol, the
Metrics bookkeeping and the spectral check -- against files / values produced by the reference itself
(tests/golden/make_golden.py::gen_pipeline), bit for bit.

Device work goes through the C ABI on both backends of test_parity.py (`emu` here, `cuda` on the B200 box).
"""
import argparse
import importlib
import os
import random

import numpy as np
import pytest
import torch

from conftest import assert_bitwise, golden_names, load_golden, rel_err
import emu_backend
from test_parity import Backend, BACKENDS


@pytest.fixture(params=BACKENDS)
def be(request):
    b = Backend(request.param)
    yield b
    emu_backend.use_cuda_library(b.npde)


geo = importlib.import_module("neural-pde-solver_b200.geometries")


# ---------------------------------------------------------------- host-side generators (no device) --
def test_filled_disc_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    for n in (17, 33, 65, 257, 1025):
        for _ in range(25):
            r = int(rng.randint(0, n))
            cx, cy = (int(v) for v in rng.randint(-5, n + 5, 2))
            want = cv2.circle(np.zeros((n, n), np.uint8), (cx, cy), r, 255, -1) / 255
            assert np.array_equal(geo.filled_disc(n, cx, cy, r), want), (n, cx, cy, r)


@pytest.mark.parametrize("tag", ["Lshape33", "cylinders33_cold", "cylinders65_cold"])
def test_geometry_generators_reproduce_reference_draws(tag):
    """get_geometry draws from np.random exactly like utils/geometries.py: the stored frames of a seeded reference
    run hold bc_values / bc_mask (channels 2, 3) and, for a cold start, x0 (channel 0)."""
    g = load_golden("pipeline_" + tag)
    N, B, n_runs = (int(v) for v in g["args"][:3])
    np.random.seed(666)
    for run in range(n_runs):
        x, v, m = geo.get_geometry(str(g["geometry"]), N, B, 1)
        frames = g["frames_%04d" % run]
        assert_bitwise(v * 100, frames[:, 2], "bc values")
        assert_bitwise(m, frames[:, 3], "bc mask")
        assert_bitwise(x.astype(np.float32).astype(np.float64) * 100, frames[:, 0], "x0")
    assert set(np.unique(m)) <= {0.0, 1.0}


def test_unknown_geometry_raises():
    with pytest.raises(NotImplementedError):
        geo.get_geometry("torus", 17, 1, 1)


def test_metrics_vs_reference():
    npde = importlib.import_module("neural-pde-solver_b200")
    g = load_golden("metrics_case")
    m = npde.utils.Metrics(scale=2.5, error_threshold=0.05, verbose=False)
    m.update({"model errors": torch.from_numpy(g["model_err"]), "Jacobi errors": torch.from_numpy(g["jac_err"])})
    m.update({"model errors": g["model_err"][:4] * np.float32(0.5)})
    invalid = m.invalid
    res = m.get_results()
    got = np.array([res["Jacobi"], res["model"], res["ratio"], invalid], np.float64)
    assert_bitwise(got, g["results"], "Metrics.get_results")


# ------------------------------------------------------------------- dataset pipeline (device loops) --
def _gen_opt(root, g, device):
    N, B, n_runs, poisson, use_model = (int(v) for v in g["args"])
    return argparse.Namespace(save_dir=root, batch_size=B, save_every=1, n_frames=1, n_runs=n_runs, image_size=N,
                              max_temp=100, poisson=poisson, geometry=str(g["geometry"]), use_model=use_model,
                              device=str(device), init_ckpt="")


# warm-started cases run 400 V-cycles per run (thousands of launches): too slow for the host emulator, GPU only
GPU_ONLY_CASES = ("square17", "poisson33", "Lshape33", "cylinders65_cold")
PIPELINE_CASES = ["square17_cold", "poisson17_cold", "cylinders33_cold", "square17", "poisson33", "Lshape33",
                  "cylinders65_cold"]


@pytest.mark.parametrize("tag", PIPELINE_CASES)
def test_generation_writes_the_reference_files(be, tag, tmp_path, capsys):
    """generation.py end to end: seeded draws, 400 warm-start V-cycles, Jacobi to fd_error < 1e-5 checked every
    100 iterations, rescaling, float64 x max_temp files -- identical bytes to the reference's own run; then the
    readers (HeatDataset / HeatGeometryDataset / runtime.get_data) on those files."""
    if tag in GPU_ONLY_CASES and be.kind == "emu":
        pytest.skip("400 warm-start V-cycles / thousands of sweeps: GPU only")
    g = load_golden("pipeline_" + tag)
    gen, data, rt = be.npde.generation, be.npde.data, be.npde.runtime
    opt = _gen_opt(str(tmp_path), g, be.device)
    np.random.seed(666)
    iters = gen.generate_square(opt) if opt.geometry == "square" else gen.generate_geometry(opt)
    capsys.readouterr()
    assert len(iters) == opt.n_runs and all(i % 100 == 0 and i <= gen.MAX_ITERS for i in iters)
    for run in range(opt.n_runs):
        got = np.load(os.path.join(opt.save_dir, "frames", "%04d.npy" % run))
        assert got.dtype == np.float64
        assert_bitwise(got, g["frames_%04d" % run], "%s frames %d" % (tag, run))
    if opt.geometry == "square":
        assert_bitwise(np.load(os.path.join(opt.save_dir, "bc.npy")), g["bc"], "bc.npy")
    saved = np.load(os.path.join(opt.save_dir, "opt.npy"), allow_pickle=True).item()
    assert saved.image_size == opt.image_size and saved.geometry == opt.geometry

    # readers
    random.seed(5)
    if opt.geometry == "square":
        dset = data.HeatDataset(opt.save_dir, bool(opt.poisson), 100, 0, opt.poisson)
    else:
        dset = data.HeatGeometryDataset(opt.save_dir, True, 100, 0)
    assert len(dset) == int(g["dset_len"])
    for idx in range(min(len(dset), 3)):
        sample = dset[idx]
        keys = sorted(k[len("sample%d_" % idx):] for k in g if k.startswith("sample%d_" % idx))
        assert sorted(sample) == keys
        for k in keys:
            assert_bitwise(sample[k].numpy(), g["sample%d_%s" % (idx, k)], "dataset sample %d %s" % (idx, k))
    got = rt.get_data(opt.geometry, 5, opt.save_dir, 100)
    for k in ("bc", "x", "final"):
        assert_bitwise(got[k].numpy(), g["runtime_" + k], "runtime.get_data " + k)


def test_get_solution_matches_oracle_iteration_count(be, oracle):
    """get_solution = Jacobi until max_b fd_error < 1e-5, looked at every 100 iterations: same stopping iteration
    and same field as the oracle run with the reference's loop (generation.py:72-99)."""
    rng = np.random.RandomState(3)
    B, N = 3, 17
    bc = rng.rand(B, 4).astype(np.float32)
    x = oracle.set_boundary(rng.rand(B, N, N).astype(np.float32), bc)
    ref, n_ref = x, 0
    if oracle.fd_error(ref, bc, None).max() >= 1e-5:
        for i in range(8000):
            ref = oracle.fd_step(ref, bc, None)
            n_ref = i + 1
            if (i + 1) % 100 == 0 and oracle.fd_error(ref, bc, None).max() < 1e-5:
                break
    sol, n = be.npde.generation.get_solution(be.t(x), be.t(bc), None, verbose=False)
    assert n == n_ref and n > 0
    assert_bitwise(sol, ref, "solution")


def test_heat_source_properties(be):
    """get_heat_source (generation.py:101-115): negative Gaussian scaled by U(3,3.5)/N^2, zero on the ring.
    (Its exact values are covered by channel 2 of the Poisson frames in the pipeline test above.)"""
    np.random.seed(1)
    f = be.npde.generation.get_heat_source(33, 5, be.device).cpu().numpy()
    assert f.shape == (5, 33, 33) and f.dtype == np.float32
    assert (f[:, 0, :] == 0).all() and (f[:, -1, :] == 0).all() and (f[:, :, 0] == 0).all() and (f[:, :, -1] == 0).all()
    assert (f <= 0).all() and (f.reshape(5, -1).min(axis=1) >= -3.5 / 33 ** 2 - 1e-9).all()


# ------------------------------------------------------------------------------- runtime.py protocol --
def _conv_model(be, n_layers=3, scale=0.05, act="clamp", geometry="square", seed=0):
    opt = argparse.Namespace(iterator="conv", activation=act, conv_n_layers=n_layers, mg_n_layers=3,
                             mg_pre_smoothing=2, mg_post_smoothing=2, geometry=geometry, is_train=False,
                             initialization="avg", n_evaluation_steps=200)
    model = be.npde.HeatModel(opt)
    rng = np.random.RandomState(seed)
    with torch.no_grad():
        for l in model.iterator.layers:
            l.weight.copy_(torch.from_numpy(((rng.rand(1, 1, 3, 3) - 0.5) * scale).astype(np.float32)))
    model.iterator.to(be.device)
    return opt, model


def test_runtime_protocol_steps_vs_oracle(be, oracle, tmp_path, capsys):
    """runtime.py:41-85 on a generated set: per-problem iterations to 1 % relative error are the oracle's
    (first index of the error curve below 0.01), one problem at a time and as one batch."""
    g = load_golden("pipeline_square17_cold")
    gen, rt = be.npde.generation, be.npde.runtime
    gopt = _gen_opt(str(tmp_path), g, be.device)
    np.random.seed(666)
    gen.generate_square(gopt)
    data = rt.get_data("square", 4, gopt.save_dir, 100)
    opt, model = _conv_model(be)
    w = np.stack([l.weight.detach().cpu().numpy()[0, 0] for l in model.iterator.layers])

    want = []
    for i in range(4):
        x = data["x"][i:i + 1].numpy().copy()
        bc = data["bc"][i:i + 1].numpy()
        gt = data["final"][i:i + 1].numpy()
        zero = x.copy()
        zero[:, 1:-1, 1:-1] = 0
        start = oracle.l2_error(zero, gt)
        x[:, 1:-1, 1:-1] = torch.mean(torch.from_numpy(bc), dim=1).view(1, 1, 1).numpy()
        steps = -1
        for k in range(opt.n_evaluation_steps + 1):
            if k > 0:
                x = oracle.conv_step(x, bc, None, w, "clamp")
            if k > 0 and np.float32(oracle.l2_error(x, gt)[0]) / np.float32(start[0]) < 0.01:
                steps = k
                break
        want.append(steps)
    times, steps = rt.runtime(opt, model, data, device=be.device, verbose=False)
    assert steps == [s for s in want if s >= 0] and len(times) == len(steps)
    assert all(t >= 0 for t in times)
    t_batch, first = rt.runtime_batched(opt, model, data, device=be.device)
    assert list(first) == want and t_batch >= 0
    capsys.readouterr()
    assert rt.report(times, 100) >= 0


# ------------------------------------------------------------------------------------- spectral check --
def test_spectral_matrices_vs_reference(be):
    """construct_matrix / calculate_eigenvalues / *_wraparound (utils/heat_utils.py:216-327): all probes in one
    batched call give the reference's probe-by-probe matrices.  Grids here are 9^2, 8^2 and 15^2 -- not 2^n+1 --
    so this also covers the stream kernels on odd sizes."""
    g = load_golden("spectral_conv2")
    opt, model = _conv_model(be, n_layers=2)
    with torch.no_grad():
        for l, w in zip(model.iterator.layers, g["w"]):
            l.weight.copy_(torch.from_numpy(w).view(1, 1, 3, 3))
    u = be.utils
    A, B = u.construct_matrix(np.array([[0.3, 0.1, 0.7, 0.5]]), 7, model.iter_step, device=be.device)
    assert_bitwise(A.numpy(), g["A"], "A")
    assert_bitwise(B.numpy(), g["B"], "B")
    # the literal one-probe-at-a-time loop of the reference (ATen's batch-of-one summation order): 1e-5 relative
    assert rel_err(A.numpy(), g["A_single"]) < 1e-5 and rel_err(B.numpy(), g["B_single"]) < 1e-5
    w, _ = u.calculate_eigenvalues(model, image_size=6)
    np.testing.assert_allclose(np.sort(np.abs(w)), g["eig_abs_sorted"], rtol=1e-5, atol=1e-6)  # LAPACK on equal input
    assert model.get_activation() == "clamp" and model.iterator.is_bc_mask is False
    model.change_activation("none")
    T = u.construct_matrix_wraparound(5, u.fd_step, device=be.device)
    H = u.construct_matrix_wraparound(5, model.H)
    model.change_activation("clamp")
    assert_bitwise(T.astype(np.float32), g["T_wrap"], "T wraparound")
    assert_bitwise(H.astype(np.float32), g["H_wrap"], "H wraparound")
    assert rel_err(T, g["T_wrap_single"]) < 1e-5 and rel_err(H, g["H_wrap_single"]) < 1e-5
    ww, _ = u.calculate_eigenvalues_wraparound(model, 5)
    np.testing.assert_allclose(np.sort(np.abs(ww)), g["eigw_abs_sorted"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(u.spectral_radius(H.dot(T) + T - H), float(g["rho"]), rtol=1e-5)
    assert np.sort(np.abs(w))[-1] == pytest.approx(1.0, abs=1e-5)  # the affine row; all others must be < 1
    assert np.sort(np.abs(w))[-2] < 1.0


# ------------------------------------------------------- comparison solvers of get_iterator (8f.4) --
@pytest.mark.parametrize("name", golden_names("cg_"))
def test_conjugate_gradient_vs_reference_and_oracle(be, oracle, name):
    """npde_cg_step: n CG iterations with device-resident alpha / beta.  fp32 dot products are summed in a
    different (deterministic) association than torch.sum: 1e-5 relative against reference and oracle."""
    g = load_golden(name)
    n_iters = int(g["meta"][0])
    it = be.npde.ConjugateGradient(n_iters)
    x, bc = be.t(g["x"]), be.t(g["bc"])
    y = it(x, bc, None)
    assert rel_err(be.n(y), g["y"]) <= 1e-5
    assert rel_err(be.n(y), oracle.cg_step(g["x"], n_iters)) <= 1e-5
    y2 = it(y, bc, None)
    assert rel_err(be.n(y2), g["y2"]) <= 1e-5
    assert_bitwise(be.n(it(x, bc, None)), be.n(y), "deterministic")
    assert_bitwise(be.n(x), g["x"], "input untouched")
    assert_bitwise(be.n(y)[:, :, 0], g["x"][:, :, 0], "ring kept")
    with pytest.raises(RuntimeError):
        it(x, bc, torch.zeros_like(x))
    # iterate() contract used by calculate_errors: fd rows from the device reductions
    yk, _, rows = it.iterate(x, bc, None, 2, want_fd=True)
    assert_bitwise(be.n(yk), be.n(y2), "iterate(2) = two forwards")
    assert rows.shape == (3, x.shape[0])


@pytest.mark.parametrize("name", golden_names("mgres_"))
def test_multigrid_residual_vs_reference(be, name):
    g = load_golden(name)
    L, pre, post = (int(v) for v in g["meta"])
    it = be.npde.MultigridResidualIterator(L, pre, post)
    it.is_bc_mask = g["bc"].ndim == 4
    x = be.t(g["x"])
    assert_bitwise(be.n(it(x, be.t(g["bc"]), be.t(g.get("f")))), g["y"], "MultigridResidualIterator.forward")
    assert_bitwise(be.n(x), g["x"], "input untouched")


def test_get_iterator_cg_contract(be):
    rows = np.load(os.path.join(os.path.dirname(__file__), "golden", "get_iterator_contract_cg.npy"),
                   allow_pickle=True)
    for itname, geom, name, cmp_name, ratio, is_train, is_mask, n_ops in rows:
        opt = argparse.Namespace(iterator=itname, activation="clamp", conv_n_layers=3, mg_n_layers=3,
                                 mg_pre_smoothing=2, mg_post_smoothing=2, cg_n_iters=4, geometry=geom, is_train=True)
        it, cmp_, r, tr = be.npde.get_iterator(opt)
        assert it.name() == name and cmp_.name() == cmp_name
        assert float(r) == float(ratio) and bool(tr) == bool(is_train)
        assert bool(it.is_bc_mask) == bool(is_mask) and float(it.n_operations) == float(n_ops)


# ---------------------------------------------------------------------------- test.py evaluation driver --
def test_evaluation_driver_vs_oracle(be, oracle, tmp_path, capsys):
    """evaluation.test (test.py:96-123) on a generated set: iterations-to-1 % of a Conv3 model and of Jacobi per
    problem, averaged by Metrics, equal the same quantities computed with the oracle; the eigenvalue check runs
    through the batched probes and finds a contraction."""
    g = load_golden("pipeline_square17_cold")
    gen, data = be.npde.generation, be.npde.data
    gopt = _gen_opt(str(tmp_path), g, be.device)
    np.random.seed(666)
    gen.generate_square(gopt)
    opt, model = _conv_model(be)
    for k, v in dict(dset_path=gopt.save_dir, max_temp=100, data_limit=0, poisson=0, batch_size=4, n_workers=0,
                     log_every=10, n_evaluation_steps=120).items():
        setattr(opt, k, v)
    loader = data.get_data_loader(opt)           # is_train False: the test split (run 1), no shuffling
    results, spectra = be.npde.evaluation.test(opt, model, loader)
    capsys.readouterr()

    w = np.stack([l.weight.detach().cpu().numpy()[0, 0] for l in model.iterator.layers])
    dset = data.HeatDataset(gopt.save_dir, False, 100, 0, 0)
    jac_steps, model_steps = [], []
    for idx in range(len(dset)):
        s = dset[idx]
        x, bc, gt = (s[k].numpy()[None].copy() for k in ("x", "bc", "final"))
        x[:, 1:-1, 1:-1] = torch.mean(torch.from_numpy(bc), dim=1).view(1, 1, 1).numpy()
        start = oracle.l2_error(x, gt)           # evaluate() normalises by the error of the initialised field

        def first(stepper):
            y = x
            for k in range(1, opt.n_evaluation_steps + 1):
                y = stepper(y)
                if np.float32(oracle.l2_error(y, gt)[0]) / np.float32(start[0]) < 0.01:
                    return k
            return None
        jac_steps.append(first(lambda y: oracle.fd_step(y, bc, None)))
        model_steps.append(first(lambda y: oracle.conv_step(y, bc, None, w, "clamp")))
    ok = [i for i in range(len(dset)) if jac_steps[i] is not None and model_steps[i] is not None]
    assert ok, "no problem reaches 1 % within the evaluation steps"
    assert results["Jacobi"] == pytest.approx(np.mean([jac_steps[i] for i in ok]))
    assert results["model"] == pytest.approx(np.mean([model_steps[i] for i in ok]))
    assert results["ratio"] == pytest.approx(np.mean([model_steps[i] / jac_steps[i] for i in ok]))
    w_model, w_h, w_fd = spectra
    assert w_model[-1] == pytest.approx(1.0, abs=1e-5) and w_model[-2] < 1.0 and w_fd[-2] < 1.0
    assert opt.initialization == "avg"
