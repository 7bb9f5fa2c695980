"""Checks of the NL2SOL restatement (oracle/nl2sol_port.cpp) against independent numerics.

The reference ships no solver fixtures, so these are the pins the
oracle has: numpy/scipy linear algebra for the building blocks, scipy/MINPACK on identical
data for whole fits, and the mcfit.py probe numbers recorded in BASELINE.md section 2.
"""
import ctypes as C

import numpy as np
import pytest
from scipy.optimize import least_squares

import datasets

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return a.ctypes.data_as(_dp)


def _unpack_upper_by_cols(rd, jfact, p):
    R = np.zeros((p, p))
    for k in range(p):
        R[:k, k] = jfact[:k, k]
        R[k, k] = rd[k]
    return R


def test_v2norm_dotprd(oracle):
    L = oracle.lib()
    L.fmc_oracle_v2norm.restype = C.c_double
    L.fmc_oracle_v2norm.argtypes = [C.c_int, _dp]
    L.fmc_oracle_dotprd.restype = C.c_double
    L.fmc_oracle_dotprd.argtypes = [C.c_int, _dp, _dp]
    rng = np.random.default_rng(0)
    for scale in (1e-200, 1e-3, 1.0, 1e150):
        x = rng.normal(size=57) * scale
        y = rng.normal(size=57)
        assert np.isclose(L.fmc_oracle_v2norm(57, _d(x)), np.linalg.norm(x / scale) * scale, rtol=1e-14)
        assert np.isclose(L.fmc_oracle_dotprd(57, _d(x), _d(y)), np.dot(x / scale, y) * scale, rtol=1e-12)
    z = np.zeros(5)
    assert L.fmc_oracle_v2norm(5, _d(z)) == 0.0 and L.fmc_oracle_v2norm(0, _d(z)) == 0.0


@pytest.mark.parametrize("n,p,seed", [(40, 3, 1), (200, 4, 2), (300, 10, 3), (5, 5, 4)])
def test_qrfact_matches_numpy(oracle, n, p, seed):
    L = oracle.lib()
    L.fmc_oracle_qrfact.argtypes = [C.c_int, C.c_int, _dp, _dp, _ip, _dp]
    rng = np.random.default_rng(seed)
    J = rng.normal(size=(n, p)) * np.logspace(0, 3, p)[None, :]
    r = rng.normal(size=n)
    a = np.asfortranarray(J.copy())
    alpha = np.zeros(p)
    ipiv = np.zeros(p, dtype=np.int32)
    qtr = r.copy()
    ierr = L.fmc_oracle_qrfact(n, p, _d(a), _d(alpha), ipiv.ctypes.data_as(_ip), _d(qtr))
    assert ierr == 0
    perm = ipiv - 1
    assert sorted(perm) == list(range(p))
    R = _unpack_upper_by_cols(alpha, a, p)
    JP = J[:, perm]
    assert np.allclose(R.T @ R, JP.T @ JP, rtol=1e-11, atol=1e-9 * np.abs(JP.T @ JP).max())
    # Q**T r: first p components solve R**T q = (JP)**T r; the norm is preserved
    assert np.allclose(R.T @ qtr[:p], JP.T @ r, rtol=1e-10, atol=1e-9)
    assert np.isclose(np.linalg.norm(qtr), np.linalg.norm(r), rtol=1e-13)
    # column pivoting: |diag R| non-increasing
    assert (np.diff(np.abs(alpha)) <= 1e-12 * np.abs(alpha).max()).all()


def test_cholesky_and_triangular_solves(oracle):
    L = oracle.lib()
    L.fmc_oracle_lsqrt.argtypes = [C.c_int, _dp, _dp]
    L.fmc_oracle_livmul.argtypes = [C.c_int, _dp, _dp, _dp]
    L.fmc_oracle_litvmu.argtypes = [C.c_int, _dp, _dp, _dp]
    rng = np.random.default_rng(5)
    p = 6
    A = rng.normal(size=(p, p))
    A = A @ A.T + p * np.eye(p)
    tri = np.tril_indices(p)
    a = np.ascontiguousarray(A[tri])  # lower triangle by rows
    l = np.zeros_like(a)
    assert L.fmc_oracle_lsqrt(p, _d(l), _d(a)) == 0
    Lm = np.zeros((p, p))
    Lm[tri] = l
    assert np.allclose(Lm, np.linalg.cholesky(A), rtol=1e-12)
    y = rng.normal(size=p)
    x = np.zeros(p)
    L.fmc_oracle_livmul(p, _d(x), _d(l), _d(y))
    assert np.allclose(Lm @ x, y, rtol=1e-12)
    L.fmc_oracle_litvmu(p, _d(x), _d(l), _d(y))
    assert np.allclose(Lm.T @ x, y, rtol=1e-12)
    a[2] = -1.0  # (2,2) entry negative -> not positive definite at row 2
    assert L.fmc_oracle_lsqrt(p, _d(l), _d(a)) == 2


def _tr_step(oracle, model, J, r, d, radius, S=None, stppar0=0.0, rad0=0.0):
    L = oracle.lib()
    L.fmc_oracle_tr_step.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_double, C.c_double,
                                     C.c_double, _dp, _dp, _ip]
    n, p = J.shape
    jf = np.asfortranarray(J.copy())
    step, vout, ka = np.zeros(p), np.zeros(9), C.c_int(0)
    sp = None if S is None else np.ascontiguousarray(S[np.tril_indices(p)])
    L.fmc_oracle_tr_step(model, n, p, _d(jf), _d(np.ascontiguousarray(r)), _d(np.ascontiguousarray(d)),
                         None if sp is None else _d(sp), radius, stppar0, rad0, _d(step), _d(vout), C.byref(ka))
    return step, vout, ka.value


@pytest.mark.parametrize("model", [1, 2])
def test_trust_region_steps(oracle, model):
    rng = np.random.default_rng(11)
    n, p = 120, 4
    J = rng.normal(size=(n, p)) * np.array([1.0, 5.0, 0.2, 30.0])
    r = rng.normal(size=n)
    d = np.sqrt((J * J).sum(0))
    g = J.T @ r
    gn = -np.linalg.lstsq(J, r, rcond=None)[0]
    # (a) huge radius: the Gauss-Newton / Newton step, stppar = 0
    step, v, ka = _tr_step(oracle, model, J, r, d, 1e6)
    assert np.allclose(step, gn, rtol=1e-10)
    assert v[4] == 0.0 and ka == 0
    assert np.isclose(v[1], np.linalg.norm(d * gn), rtol=1e-12)           # dstnrm
    assert np.isclose(v[3], g @ step, rtol=1e-12)                          # gtstep
    assert np.isclose(v[5], 0.5 * (np.linalg.norm(r) ** 2 - np.linalg.norm(r + J @ gn) ** 2), rtol=1e-9)  # nreduc
    # (b) small radius: (J^T J + alpha D^2) step = -g with ||D step|| within [0.9, 1.1] radius
    radius = 0.3 * np.linalg.norm(d * gn)
    step, v, ka = _tr_step(oracle, model, J, r, d, radius)
    alpha = v[4]
    assert alpha > 0 and ka >= 1
    assert 0.9 * radius <= np.linalg.norm(d * step) <= 1.1 * radius
    assert np.allclose((J.T @ J + alpha * np.diag(d * d)) @ step, -g, rtol=1e-9, atol=1e-9 * np.abs(g).max())
    pred = -(g @ step) - 0.5 * step @ (J.T @ J) @ step
    assert np.isclose(v[6], pred, rtol=1e-9)                               # preduc


def test_gqt_step_indefinite_hessian(oracle):
    rng = np.random.default_rng(12)
    n, p = 60, 3
    J = rng.normal(size=(n, p))
    r = rng.normal(size=n)
    d = np.ones(p)
    S = -3.0 * (J.T @ J)  # J^T J + S negative definite
    radius = 0.5
    step, v, ka = _tr_step(oracle, 2, J, r, d, radius, S=S)
    H = J.T @ J + S
    g = J.T @ r
    alpha = abs(v[4])
    assert 0.9 * radius <= np.linalg.norm(step) <= 1.1 * radius
    assert np.linalg.eigvalsh(H + alpha * np.eye(p)).min() > -1e-8
    assert np.allclose((H + alpha * np.eye(p)) @ step, -g, rtol=1e-6, atol=1e-8)
    assert v[5] == 0.0  # nreduc = 0 for an indefinite H (toms573.f90:2696)


def _scipy_fit(model, x, y, err, p0):
    f = lambda q: (y - datasets.model_y(model, q, x)) / err
    return least_squares(f, p0, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, x_scale="jac").x


@pytest.mark.parametrize("name", ["c1", "c2p3", "c2", "c4", "c5mild"])
def test_fit_matches_scipy(oracle, name):
    model, x, y, err = datasets.dataset(name)
    info = oracle.model_info(model)
    p, counters = oracle.fit_data(model, x, y, err, info["params_init"])
    assert counters[0][0] in (3, 4, 5, 6) and counters[1][0] in (3, 4, 5, 6)
    ref = _scipy_fit(model, x, y, err, p)  # start scipy at the answer: both polish the same minimum
    # forward differences bias the stationary point by ~sqrt(eps) * curvature: 1e-7 is the honest bound
    assert np.allclose(p, ref, rtol=2e-7, atol=0), (p, ref)
    # gradient of chi^2 at the oracle's answer is negligible relative to its scale
    r = oracle.madr(model, x, y, err, p)
    eps = 1e-6
    Jn = np.array([(oracle.madr(model, x, y, err, p + eps * np.abs(p[k]) * np.eye(len(p))[k]) - r) / (eps * np.abs(p[k]))
                   for k in range(len(p))]).T
    g = Jn.T @ r
    assert (np.abs(g) <= 1e-4 * np.sqrt((Jn * Jn).sum(0)) * np.linalg.norm(r)).all()


def test_probe_numbers_from_baseline_md(oracle):
    """mcfit.py (scipy curve_fit, default tolerances) on the probe file."""
    model, x, y, err = datasets.dataset("c1")
    p, _ = oracle.fit_data(0, x, y, err, [0.0, 2.0, 1.0])
    want = np.array([0.5039360403603528, 1.960438556490636, 1.2822067222414264])
    assert np.allclose(p, want, rtol=5e-8)
    r = oracle.madr(0, x, y, err, p)
    assert np.isclose((r * r).sum() / (40 - 3), 1.3016073130752277, rtol=1e-12)
    assert np.isclose(oracle.derived_properties(0, p)[0], 2.4643745968509885, rtol=5e-9)


def test_dead_work_does_not_change_parameters(oracle):
    """covclc restores x and r exactly (toms573.f90:2142-2180), and the per-call Jacobian
    copy is a pure copy: both can be dropped without changing a single bit of the fit."""
    model, x, y, err = datasets.dataset("c2")
    p0 = oracle.model_info(model)["params_init"]
    lean, c_lean = oracle.fit_data(model, x, y, err, p0, flags=oracle.LSVMIN_PER_FIT)
    full, c_full = oracle.fit_data(model, x, y, err, p0,
                                   flags=oracle.LSVMIN_PER_FIT | oracle.COVARIANCE | oracle.COPY_JAC)
    assert np.array_equal(lean, full)
    pcount = len(p0)
    assert c_full[0][4] == pcount + pcount * (pcount + 1) // 2 and c_lean[0][4] == 0  # nfcov
    assert c_full[0][1] - c_full[0][4] == c_lean[0][1]


def test_bootstrap_loop_sums_and_injection(oracle):
    model, x, y, err = datasets.dataset("c1")
    start, _ = oracle.fit_data(model, x, y, err, [0.0, 2.0, 1.0])
    rng = np.random.default_rng(7)
    N = 64
    z = rng.normal(size=(N, len(x)))
    out = oracle.bootstrap(model, x, y, err, start, N, z=z)
    for i in (0, 17, 63):
        pi, ci = oracle.fit_data(model, x, y + z[i] * err, err, start)
        assert np.array_equal(pi, out["params"][i])
        assert np.array_equal(ci, out["counters"][i])
    assert np.allclose(out["sum_p"], out["params"].sum(0), rtol=1e-13)
    assert np.allclose(out["sum_p2"], (out["params"] ** 2).sum(0), rtol=1e-13)
    assert np.allclose(out["props"][:, 0], out["params"][:, 0] + out["params"][:, 1], rtol=1e-15)
    assert out["rc_count"].sum() == N
    par = oracle.bootstrap(model, x, y, err, start, N, z=z, nthreads=4)
    assert np.array_equal(par["params"], out["params"])
    mean, sd = oracle.finalise_moments(N, out["sum_p"], out["sum_p2"])
    assert np.allclose(mean, out["params"].mean(0)) and np.allclose(sd, out["params"].std(0, ddof=1), rtol=1e-8)


def test_serial_ranlux_stream_feeds_the_loop(oracle):
    """offset_data draws rang(ndata) per resample from the one global stream (bootstrap.f90:205-208)."""
    model, x, y, err = datasets.dataset("c1")
    start, _ = oracle.fit_data(model, x, y, err, [0.0, 2.0, 1.0])
    oracle.initialise_rng()
    out = oracle.bootstrap(model, x, y, err, start, 5)
    oracle.initialise_rng()
    z = np.array([oracle.rang(len(x)) for _ in range(5)])
    inj = oracle.bootstrap(model, x, y, err, start, 5, z=z)
    assert np.array_equal(out["params"], inj["params"])


def test_ensemble_agrees_with_scipy_on_same_resamples(oracle):
    """The mcfit.py cross-check (mcfit.py:90-115) on identical resampled data sets."""
    model, x, y, err = datasets.dataset("c1")
    start, _ = oracle.fit_data(model, x, y, err, [0.0, 2.0, 1.0])
    rng = np.random.default_rng(99)
    N = 200
    z = rng.normal(size=(N, len(x)))
    out = oracle.bootstrap(model, x, y, err, start, N, z=z)
    ref = np.array([_scipy_fit(model, x, y + z[i] * err, err, out["params"][i]) for i in range(N)])
    assert np.allclose(out["params"], ref, rtol=5e-7)
    # survey's MC numbers (unseeded 2000 samples): statistical agreement only
    mean, sd = out["params"].mean(0), out["params"].std(0, ddof=1)
    want_m = np.array([0.50388, 1.96112, 1.28296])
    want_s = np.array([0.01247, 0.03474, 0.04703])
    assert (np.abs(mean - want_m) < 3 * want_s * np.sqrt(1 / N + 1 / 2000)).all()
    assert (np.abs(sd - want_s) < 4 * want_s * np.sqrt(0.5 / N + 0.5 / 2000)).all()
