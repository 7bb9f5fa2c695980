"""Analytic-Jacobian path: nl2sol + madj, the switch the reference's
README (35-42) describes, against the oracle's restatement of nl2sol (toms573.f90:35-587) with
hand-derived madj formulas.

The product forms its Jacobian rows by dual numbers from the same generic model_y that serves
the forward-difference path (csrc/dual.cuh), or from a plug-in's own model_grad; the oracle's
madj is derived by hand.  The two are independent, so their agreement checks both.  CPU part:
the product's solver and row code compiled for the host (tests/host_emul).  GPU part: the CUDA
path through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import datasets

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL_DIR = os.path.join(HERE, "host_emul")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

REL_MEDIAN, REL_MAX, SIGMA_MAX = 1e-8, 5e-7, 1e-4     # the gates of tests/test_gpu_parity.py


def _d(a):
    return a.ctypes.data_as(_dp)


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL_DIR, "libfmc_emul.so")
    src = os.path.join(EMUL_DIR, "emul.cpp")
    csrc = os.path.join(os.path.dirname(HERE), "fitter_mc_b200", "csrc")
    deps = [src] + [os.path.join(dp, f) for dp, _, fs in os.walk(csrc) for f in fs if f.endswith((".cuh", ".inc"))]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-x", "c++", "-std=c++17", "-fPIC", "-shared", "-w", "-o", so, src])
    L = C.CDLL(so)
    L.emul_bootstrap_jac.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_longlong, _dp, _dp, _dp,
                                     _ip, _ip]
    L.emul_row_grad.argtypes = [C.c_int, _dp, C.c_double, _dp]
    return L


def _emul_run(emul, model, jac, x, y, err, start, z, nparam, nprop):
    nres = len(z)
    po, qo = np.zeros((nres, nparam)), np.zeros((nres, max(nprop, 1)))
    co, an = np.zeros((nres, 2, 5), dtype=np.int32), np.zeros(nres, dtype=np.int32)
    rc = emul.emul_bootstrap_jac(model, jac, len(x), _d(x), _d(y), _d(err), _d(np.ascontiguousarray(start)), nres,
                                 _d(z), _d(po), _d(qo), co.ctypes.data_as(_ip), an.ctypes.data_as(_ip))
    assert rc == 0
    return po, qo, co, an


def _gates(name, got, want):
    rel = (np.abs(got - want) / np.abs(want)).max(1)
    sig = (np.abs(got - want) / want.std(0)).max()
    print("\n[analytic %s] n=%d rel dev: median %.2e p99 %.2e max %.2e ; max %.2e sigma"
          % (name, len(got), np.median(rel), np.quantile(rel, 0.99), rel.max(), sig))
    assert np.median(rel) <= REL_MEDIAN and rel.max() <= REL_MAX and sig <= SIGMA_MAX
    return rel


@pytest.mark.parametrize("model", [0, 1, 2, 3])
def test_dual_number_rows_equal_the_hand_derived_madj(emul, oracle, model):
    rng = np.random.default_rng(40 + model)
    p0 = np.array(datasets.TRUTH[model])
    for _ in range(20):
        p = p0 * (1.0 + 0.2 * rng.uniform(-1, 1, len(p0)))
        xs = np.sort(rng.uniform(0.15, 9.0, 12))
        want = -oracle.madj(model, xs, np.ones_like(xs), p)          # d model_y / d params, [n][p]
        for i, xi in enumerate(xs):
            g = np.zeros(len(p))
            assert emul.emul_row_grad(model, _d(p), float(xi), _d(g)) == 0
            assert np.allclose(g, want[i], rtol=2e-14, atol=1e-300), (model, xi, g, want[i])


def test_oracle_madj_is_the_derivative_of_madr(oracle):
    """Pins the hand-derived formulas themselves: central differences of the residual routine."""
    for name in ("c1", "c2", "c4", "c5mild"):
        model, x, y, err = datasets.dataset(name)
        p0 = np.array(datasets.TRUTH[model])
        J = oracle.madj(model, x, err, p0)
        for k in range(len(p0)):
            h = 1e-6 * max(abs(p0[k]), 1.0)
            pp, pm = p0.copy(), p0.copy()
            pp[k] += h
            pm[k] -= h
            jn = (oracle.madr(model, x, y, err, pp) - oracle.madr(model, x, y, err, pm)) / (2 * h)
            assert np.abs(J[:, k] - jn).max() <= 2e-9 * np.abs(J).max()


def test_oracle_nl2sol_agrees_with_scipy_and_with_nl2sno(oracle):
    from scipy.optimize import least_squares

    for name in ("c1", "c2", "c5mild"):
        model, x, y, err = datasets.dataset(name)
        init = oracle.model_info(model)["params_init"]
        fd, _ = oracle.fit_data(model, x, y, err, init)
        an, cnt = oracle.fit_data(model, x, y, err, init, flags=oracle.LSVMIN_PER_FIT | oracle.ANALYTIC_JAC)
        assert cnt[1, 0] in (3, 4, 5, 6)
        assert np.allclose(an, fd, rtol=5e-8)                        # same minimum; FD truncation ~1e-8
        ref = least_squares(lambda p: oracle.madr(model, x, y, err, p), an, jac=lambda p: oracle.madj(model, x, err, p),
                            xtol=1e-15, ftol=1e-15, gtol=1e-15)
        assert np.allclose(an, ref.x, rtol=2e-7)
        g = oracle.madj(model, x, err, an).T @ oracle.madr(model, x, y, err, an)
        gfd = oracle.madj(model, x, err, fd).T @ oracle.madr(model, x, y, err, fd)
        assert np.abs(g).max() <= 1e-4 * np.abs(oracle.madj(model, x, err, an)).max() * np.sqrt(len(x))
        assert np.isfinite(gfd).all()


@pytest.mark.parametrize("name,nres", [("c1", 400), ("c2p3", 100), ("c2", 100), ("c5mild", 100), ("c4", 8)])
def test_emulated_analytic_path_matches_oracle_nl2sol_per_fit(emul, oracle, name, nres):
    model, x, y, err = datasets.dataset(name)
    info = oracle.model_info(model)
    flags = oracle.LSVMIN_PER_FIT | oracle.ANALYTIC_JAC
    start, _ = oracle.fit_data(model, x, y, err, info["params_init"], flags=flags)
    z = np.random.default_rng(900 + model).normal(size=(nres, len(x)))
    want = oracle.bootstrap(model, x, y, err, start, nres, z=z, flags=flags, nthreads=8)
    got, props, co, an = _emul_run(emul, model, 1, x, y, err, start, z, info["nparam"], info["nprop"])
    assert not an.any()
    _gates(name, got, want["params"])
    assert np.allclose(props[:, :info["nprop"]], want["props"], rtol=1e-6)
    # Same iteration history in the first nl2sol call of (nearly) every fit: the same state
    # machine.  The second call restarts AT the converged point, where the proposed steps are of
    # rounding size, so whether one more of them is accepted is rounding noise (normal equations
    # here, Householder QR in the reference) -- printed, not asserted.
    wc = want["counters"][:, :, :4]
    same0 = (co[:, 0, :4] == wc[:, 0]).all(1).mean()
    same1 = (co[:, 1, :4] == wc[:, 1]).all(1).mean()
    print("[analytic %s] identical (rc, nfcall, ngcall, niter): first call %.3f, second call %.3f"
          % (name, same0, same1))
    assert same0 >= 0.8


def test_analytic_and_forward_difference_paths_agree_to_fd_truncation(emul, oracle):
    model, x, y, err = datasets.dataset("c1")
    info = oracle.model_info(model)
    start, _ = oracle.fit_data(model, x, y, err, info["params_init"])
    z = np.random.default_rng(5).normal(size=(300, len(x)))
    fd, _, _, _ = _emul_run(emul, model, 0, x, y, err, start, z, 3, 1)
    an, _, _, _ = _emul_run(emul, model, 1, x, y, err, start, z, 3, 1)
    rel = (np.abs(fd - an) / np.abs(an)).max(1)
    print("\n[analytic vs fd] median %.2e max %.2e" % (np.median(rel), rel.max()))
    assert 0 < rel.max() <= 5e-7 and np.median(rel) <= 2e-8          # opt-in: results move at the 1e-8 level


def test_plugin_with_its_own_model_grad_takes_that_route(emul, oracle):
    """A double-only plug-in that ships a madj (model_grad) gives the dual-number result."""
    model, x, y, err = datasets.dataset("c1")
    start, _ = oracle.fit_data(model, x, y, err, oracle.model_info(model)["params_init"])
    z = np.random.default_rng(6).normal(size=(200, len(x)))
    dual, _, _, _ = _emul_run(emul, 0, 1, x, y, err, start, z, 3, 1)
    hand, _, _, _ = _emul_run(emul, 100, 1, x, y, err, start, z, 3, 1)
    assert np.allclose(dual, hand, rtol=1e-9)
    g1, g2, p = np.zeros(3), np.zeros(3), np.array([0.5, 2.0, 1.3])
    emul.emul_row_grad(0, _d(p), 1.7, _d(g1))
    emul.emul_row_grad(100, _d(p), 1.7, _d(g2))
    assert np.allclose(g1, g2, rtol=1e-15)


@pytest.fixture(scope="module")
def fmc():
    import fitter_mc_b200

    fitter_mc_b200.lib()
    return fitter_mc_b200


def test_every_compiled_in_model_offers_the_analytic_mode(fmc):
    for model in range(fmc.model_count()):
        assert fmc.lib().fmc_model_has_analytic_jacobian(model) == 1
    assert fmc.lib().fmc_model_has_analytic_jacobian(99) == 0
    assert fmc.lib().fmc_set_jacobian_mode(None, 7) == -4            # FMC_ERR_BAD_ARG
    assert fmc.lib().fmc_set_jacobian_mode(None, 0) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name,nres", [("c1", 2000), ("c2", 600), ("c5mild", 400), ("c4", 24)])
def test_cuda_analytic_path_matches_oracle_nl2sol(fmc, oracle, name, nres):
    import torch

    model, x, y, err = datasets.dataset(name)
    info = oracle.model_info(model)
    flags = oracle.LSVMIN_PER_FIT | oracle.ANALYTIC_JAC
    start, _ = oracle.fit_data(model, x, y, err, info["params_init"], flags=flags)
    z = np.random.default_rng(1900 + model).normal(size=(nres, len(x)))
    want = oracle.bootstrap(model, x, y, err, start, nres, z=z, flags=flags, nthreads=8)
    ctx = fmc.Context(0)
    ctx.set_data(model, x, y, err)
    ctx.set_jacobian_mode("analytic")
    zd = torch.from_numpy(z).cuda()
    ctx.launch(start, nres, z_inject_dev=zd.data_ptr(), keep=True)
    got = ctx.finish()
    assert ctx.anomalies()[2] == 0
    _gates("cuda/" + name, got["params"], want["params"])
    assert got["rc_count"][3:7].sum() >= 0.9 * nres                  # converged like the oracle's fits
    # and the forward-difference mode on the same context still is the default path
    ctx.set_jacobian_mode("forward-difference")
    ctx.launch(start, nres, z_inject_dev=zd.data_ptr(), keep=True)
    fd = ctx.finish()
    rel = (np.abs(fd["params"] - got["params"]) / np.abs(got["params"])).max(1)
    print("[cuda/%s] forward-difference vs analytic: median %.2e max %.2e" % (name, np.median(rel), rel.max()))
    assert 0 < rel.max() <= 5e-6 and np.median(rel) <= 1e-7
    ctx.close()


@pytest.mark.gpu
def test_cuda_analytic_mode_through_the_one_call_entry_points(fmc, oracle):
    model, x, y, err = datasets.dataset("c1")
    info = oracle.model_info(model)
    flags = oracle.LSVMIN_PER_FIT | oracle.ANALYTIC_JAC
    want, _ = oracle.fit_data(model, x, y, err, info["params_init"], flags=flags)
    try:
        assert fmc.lib().fmc_set_jacobian_mode(None, 1) == 0
        got, rc = fmc.fit_data(x, y, err, info["params_init"], model)
        assert np.allclose(got, want, rtol=1e-8)
        a = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=5000, params_init=want, initial_fit=False)
    finally:
        fmc.lib().fmc_set_jacobian_mode(None, 0)
    b = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=5000, params_init=want, initial_fit=False)
    assert np.allclose(a["params_opt"], b["params_opt"], rtol=1e-7)  # same resamples, Jacobian differs at 1e-8
    assert not np.array_equal(a["sum_p"], b["sum_p"])
