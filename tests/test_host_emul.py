"""CPU checks of the PRODUCT's solver logic, compiled for the host (tests/host_emul/emul.cpp):
the same per-fit NL2SOL state machine (nl2_solver.cuh), per-row pass code (pass_eval.cuh,
pert.cuh) and generator (rng.cuh) that the CUDA kernels instantiate, driven one fit after
another.  It lets the algorithmic redesign (no stored Jacobian, Gram-matrix Cholesky instead
of Householder QR, speculative trial passes, perturbation-arithmetic forward differences) be
compared with the oracle in a container without a GPU.  The harness is test-only code."""
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
    L.emul_bootstrap.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_longlong, _dp, C.c_ulonglong,
                                 C.c_longlong, _dp, _dp, _ip]
    L.emul_row_model.argtypes = [C.c_int, _dp, _dp, C.c_double, _dp, _dp, _dp]
    L.emul_normals.argtypes = [C.c_ulonglong, C.c_longlong, C.c_int, _dp]
    return L


def test_philox4x32_10_known_answers(emul):
    """Random123 known-answer vectors for Philox4x32-10."""
    out = (C.c_uint * 4)()
    emul.emul_philox(0, 0, 0, 0, 0, 0, out)
    assert list(out) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    emul.emul_philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, out)
    assert list(out) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    emul.emul_philox(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0, out)
    assert list(out) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_counter_normals_are_standard_normal_and_keyed_by_id(emul):
    n = 2_000_000
    z = np.zeros(n)
    emul.emul_normals(31415, 7, n, _d(z))
    assert abs(z.mean()) < 4 / np.sqrt(n) and abs(z.std() - 1) < 4 / np.sqrt(2 * n)
    assert abs((z ** 3).mean()) < 4 * np.sqrt(15 / n) and abs((z ** 4).mean() - 3) < 4 * np.sqrt(96 / n)
    z2 = np.zeros(1000)
    emul.emul_normals(31415, 7, 1000, _d(z2))
    assert np.array_equal(z2, z[:1000])                       # reproducible
    emul.emul_normals(31415, 8, 1000, _d(z2))
    assert abs(np.corrcoef(z2, z[:1000])[0, 1]) < 0.15         # another resample id: another stream
    emul.emul_normals(31416, 7, 1000, _d(z2))
    assert abs(np.corrcoef(z2, z[:1000])[0, 1]) < 0.15         # another seed: another stream


def _mp_model(mp, model, p, x):
    p = [mp.mpf(float(v)) for v in p]
    x = mp.mpf(float(x))
    if model == 0:
        return p[0] + p[1] * mp.exp(-p[2] * x)
    if model == 1:
        return p[0] + p[1] * mp.exp(-p[2] * x) + p[3] * x
    if model == 2:
        y = p[0]
        for k in range(3):
            u = (x - p[2 + 3 * k]) / p[3 + 3 * k]
            y += p[1 + 3 * k] * mp.exp(-0.5 * u * u)
        return y
    return p[0] * mp.exp(-p[1] * x) + p[2] * x ** (-p[3])


@pytest.mark.parametrize("model", [0, 1, 2, 3])
def test_perturbation_arithmetic_is_the_forward_difference(emul, model):
    """pert.cuh: the engine's dm[k] equals the exact numerator m(x) - m(x + h_k e_k) of nl2sno's
    forward difference (toms573.f90:714-733) to ~7e-9 relative -- the rounding of x_k + h_k that
    the reference itself commits -- and is far closer to it than the plain double subtraction."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 80
    assert emul.emul_uses_pert(model) == 1
    rng = np.random.default_rng(model)
    truth = datasets.TRUTH[model]
    P = len(truth)
    worst_engine, worst_plain = 0.0, 0.0
    for _ in range(60):
        x = truth * (1 + 0.1 * rng.normal(size=P))
        h = 1.4901161193847656e-08 * np.maximum(np.abs(x), 1.0 / rng.uniform(0.5, 500, size=P))
        xi = float(rng.uniform(0.2, 9.5))
        m0, de, dpl = C.c_double(), np.zeros(P), np.zeros(P)
        emul.emul_row_model(model, _d(x), _d(h), xi, C.byref(m0), _d(de), _d(dpl))
        base = _mp_model(mp, model, x, xi)
        # (model 3 forms x^(-d) as exp(-d log x_i) from the stored log x_i: the rounding of d log x_i,
        # up to |d log x| ulp, shows in the value -- a few 1e-16, the same in the oracle's model)
        assert abs(mp.mpf(m0.value) - base) <= (2e-15 if model == 3 else 4e-16) * abs(base)
        scale = max(abs(_mp_model(mp, model, x, xi) - _mp_model(mp, model, x * (1 + 1e-8), xi)), mp.mpf(1e-300))
        for k in range(P):
            xs = x.copy()
            xs[k] = x[k] + h[k]
            exact = base - _mp_model(mp, model, xs, xi)
            if abs(exact) < 1e-12 * scale:
                continue  # a column that is numerically zero at this point (far Gaussian tail)
            worst_engine = max(worst_engine, float(abs(mp.mpf(de[k]) - exact) / abs(exact)))
            worst_plain = max(worst_plain, float(abs(mp.mpf(dpl[k]) - exact) / abs(exact)))
    assert worst_engine < 2e-8, worst_engine
    assert worst_plain > worst_engine


@pytest.mark.parametrize("name,nres", [("c1", 1500), ("c2p3", 200), ("c2", 200), ("c5mild", 200), ("c4", 32)])
def test_emulated_product_solver_matches_oracle_per_fit(emul, oracle, name, nres):
    model, x, y, err = datasets.dataset(name)
    info = oracle.model_info(model)
    P, Q = info["nparam"], info["nprop"]
    start, _ = oracle.fit_data(model, x, y, err, info["params_init"])
    z = np.random.default_rng(1).normal(size=(nres, len(x)))
    want = oracle.bootstrap(model, x, y, err, start, nres, z=z, nthreads=8)
    po, qo = np.zeros((nres, P)), np.zeros((nres, max(Q, 1)))
    co = np.zeros((nres, 2, 5), dtype=np.int32)
    emul.emul_bootstrap(model, len(x), _d(x), _d(y), _d(err), _d(start), nres, _d(z), 0, 0, _d(po), _d(qo),
                        co.ctypes.data_as(_ip))
    rel = (np.abs(po - want["params"]) / np.abs(want["params"])).max(1)
    sig = (np.abs(po - want["params"]) / want["params"].std(0)).max()
    print("\n[emul %s] median %.2e p99 %.2e max %.2e ; %.2e sigma" % (name, np.median(rel), np.quantile(rel, 0.99),
                                                                    rel.max(), sig))
    assert np.median(rel) <= 1e-8 and rel.max() <= 5e-7 and sig <= 1e-4
    # the first nl2sno call follows the oracle's iteration path in the vast majority of fits
    same_first_call = (co[:, 0, :4] == want["counters"][:, 0, :4]).all(1).mean()
    assert same_first_call > 0.7, same_first_call
    # a pass over the data is needed only for the start of each call and for each trial point
    assert (co[:, :, 4] == co[:, :, 1]).mean() > 0.99


def _mp_all_ops(mp, p, x):
    x = mp.mpf(float(x))
    a = mp.sqrt(p[0] * x + 1) * mp.sin(p[1] * x) - mp.cos(p[2]) / (1 + p[0] * p[0])
    b = mp.log(p[1] + x) * p[2] ** p[0] + (x + 1) ** p[1] - (p[0] + 2) ** mp.mpf(1.5)
    c = (-p[2] + 3 * x) / (x + p[1]) + 2 / p[0] - p[1] + (mp.mpf(0.5) - p[2]) * (x - p[0]) + mp.exp(-p[0] * x)
    return a + b / 4 + c


def test_every_operation_of_pert_and_dual_against_multiprecision(emul):
    """A plug-in may use + - * / exp log sqrt sin cos pow in any combination (pert.cuh, dual.cuh); the
    bundled models use only some of them.  A test model that uses all of them: the perturbation
    result equals m(x) - m(x + h e_k) and the dual-number result equals dm/dp_k, both computed with
    80 digits."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 80
    emul.emul_row_grad.argtypes = [C.c_int, _dp, C.c_double, _dp]
    rng = np.random.default_rng(101)
    worst_pert, worst_dual, worst_plain = 0.0, 0.0, 0.0
    for _ in range(200):
        x = np.array([rng.uniform(0.3, 3.0), rng.uniform(0.3, 3.0), rng.uniform(0.3, 3.0)])
        h = 1.4901161193847656e-08 * np.maximum(np.abs(x), 1.0 / rng.uniform(0.5, 50, size=3))
        xi = float(rng.uniform(0.1, 4.0))
        m0, de, dpl, gr = C.c_double(), np.zeros(3), np.zeros(3), np.zeros(3)
        assert emul.emul_row_model(101, _d(x), _d(h), xi, C.byref(m0), _d(de), _d(dpl)) == 0
        assert emul.emul_row_grad(101, _d(x), xi, _d(gr)) == 0
        pm = [mp.mpf(float(v)) for v in x]
        base = _mp_all_ops(mp, pm, xi)
        assert abs(mp.mpf(m0.value) - base) <= 2e-15 * max(abs(base), 1)
        for k in range(3):
            ps = list(pm)
            ps[k] = pm[k] + mp.mpf(float(h[k]))                    # the exact step, as pert.cuh carries it
            exact = base - _mp_all_ops(mp, ps, xi)
            worst_pert = max(worst_pert, float(abs(mp.mpf(float(de[k])) - exact) / abs(exact)))
            worst_plain = max(worst_plain, float(abs(mp.mpf(float(dpl[k])) - exact) / abs(exact)))
            deriv = mp.diff(lambda t, k=k: _mp_all_ops(mp, [t if i == k else pm[i] for i in range(3)], xi), pm[k])
            worst_dual = max(worst_dual, float(abs(mp.mpf(float(gr[k])) - deriv) / abs(deriv)))
    print("\n[all ops] worst rel. error: perturbation %.2e (plain double differences %.2e), dual numbers %.2e"
          % (worst_pert, worst_plain, worst_dual))
    assert worst_pert < 5e-13 and worst_dual < 5e-13                # cancellation between the terms of the model
    assert worst_plain > 1e3 * worst_pert                           # what the engine avoids


@pytest.mark.parametrize("name,nres,jac", [("c1", 300, 0), ("c2", 120, 0), ("c2", 60, 1), ("c5mild", 80, 0), ("c5", 60, 0),
                                          ("c4", 6, 0)])
def test_residual_first_mode_takes_the_same_decisions(emul, oracle, name, nres, jac):
    """Nl2::spec = 0 (the round schedule for many parameters): a trial point gets its residual sum alone
    first, and only an accepted step that does not end the call asks for the Jacobian pass.  The solver
    then reads the same numbers in the same order as with the Jacobian evaluated at every trial point, so
    the fits -- parameters, return codes, evaluation and iteration counts -- are identical bit for bit;
    only the number of passes differs (one more per accepted, continuing step)."""
    from test_analytic_jacobian import _emul_run

    model, x, y, err = datasets.dataset(name)
    info = oracle.model_info(model)
    start = datasets.C5_FITTED.copy() if name == "c5" else oracle.fit_data(model, x, y, err, info["params_init"])[0]
    z = np.random.default_rng(77).normal(size=(nres, len(x)))
    emul.emul_set_spec.argtypes = [C.c_int]
    emul.emul_bootstrap_jac.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_longlong, _dp, _dp, _dp,
                                        _ip, _ip]
    try:
        emul.emul_set_spec(1)
        p1, q1, c1, a1 = _emul_run(emul, model, jac, x, y, err, start, z, info["nparam"], info["nprop"])
        emul.emul_set_spec(0)
        p0, q0, c0, a0 = _emul_run(emul, model, jac, x, y, err, start, z, info["nparam"], info["nprop"])
    finally:
        emul.emul_set_spec(1)
    assert np.array_equal(p1, p0) and np.array_equal(q1, q0)
    assert np.array_equal(c1[:, :, :4], c0[:, :, :4]) and np.array_equal(a1, a0)
    extra = (c0[:, :, 4] - c1[:, :, 4]).sum(1)
    print("\n[residual first %s] %d fits identical; passes per fit %.2f -> %.2f, of which with a Jacobian %.2f -> %.2f"
          % (name, nres, c1[:, :, 4].sum(1).mean(), c0[:, :, 4].sum(1).mean(), c1[:, :, 4].sum(1).mean(),
             2 + extra.mean()))
    assert (extra >= 0).all()
