"""A fit may never spin forever (on a GPU one stuck lane hangs the whole launch), whatever the
data: degenerate, non-finite, wildly scaled or hopeless starts.  These run the product's solver
logic through the host emulation under a timeout; where the reference algorithm itself is well
defined the result is compared with the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import datasets
from test_host_emul import emul, _d, _ip  # noqa: F401  (fixture + helpers)


def _run(emul, model, x, y, err, start, z):
    x, y, err, start, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, err, start, z))
    n, N = len(x), len(z)
    P = len(start)
    po, qo = np.zeros((N, P)), np.zeros((N, 4))
    co = np.zeros((N, 2, 5), dtype=np.int32)
    an = np.zeros(N, dtype=np.int32)
    emul.emul_bootstrap2.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_longlong,
                                     C.POINTER(C.c_double), C.c_ulonglong, C.c_longlong, C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), _ip, _ip]
    emul.emul_bootstrap2(model, n, _d(x), _d(y), _d(err), _d(start), N, _d(z), 0, 0, _d(po), _d(qo),
                         co.ctypes.data_as(_ip), an.ctypes.data_as(_ip))
    return po, co, an


@pytest.mark.timeout(300)
def test_exact_data_zero_residual(emul, oracle):
    """y = model exactly, no noise: f -> 0, absolute-function convergence (return code 6)."""
    x = np.linspace(0, 5, 50)
    truth = np.array([0.5, 2.0, 1.3])
    y = datasets.model_y(0, truth, x)
    err = np.full(50, 0.05)
    po, co, an = _run(emul, 0, x, y, err, [0.4, 2.2, 1.1], np.zeros((1, 50)))
    want, cw = oracle.fit_data(0, x, y, err, [0.4, 2.2, 1.1])
    assert np.allclose(po[0], truth, rtol=1e-7) and np.allclose(po[0], want, rtol=1e-6)
    assert an[0] == 0


@pytest.mark.timeout(600)
def test_hopeless_starts_and_huge_noise_terminate(emul, oracle):
    rng = np.random.default_rng(3)
    x, y, err = datasets.c5(n=120, rel=0.8, floor=1.0)
    z = rng.normal(size=(300, 120)) * 3.0
    start = np.array([1.0, 1.0, 1.0, 1.0])
    po, co, an = _run(emul, 3, x, y, err, start, z)
    want = oracle.bootstrap(3, x, y, err, start, 300, z=z, nthreads=8)
    assert (co[:, :, 1] <= 200).all() and (co[:, :, 3] <= 150).all()      # mxfcal / mxiter respected
    assert (co[:, :, 4] <= 2 * (200 + 150)).all()
    both = np.isfinite(po).all(1) & np.isfinite(want["params"]).all(1)
    same = np.isclose(po[both], want["params"][both], rtol=1e-4).all(1)
    print("\n[robust] finite in both %d/300, agreeing to 1e-4: %.3f, anomalies %d, rc %s"
          % (both.sum(), same.mean(), np.count_nonzero(an), np.bincount(co[:, 1, 0], minlength=16)[3:11]))
    assert same.mean() > 0.8


@pytest.mark.timeout(300)
@pytest.mark.parametrize("case", ["nan_y", "inf_y", "huge_y", "constant_y", "n_equals_p", "zero_param_column"])
def test_degenerate_inputs_terminate(emul, case):
    rng = np.random.default_rng(11)
    x = np.linspace(0, 5, 40)
    err = np.full(40, 0.05)
    y = 0.5 + 2.0 * np.exp(-1.3 * x) + rng.normal(0, err)
    start = np.array([0.5, 2.0, 1.3])
    z = rng.normal(size=(8, 40))
    if case == "nan_y":
        y = y.copy(); y[7] = np.nan
    elif case == "inf_y":
        y = y.copy(); y[7] = np.inf
    elif case == "huge_y":
        y = y * 1e200
    elif case == "constant_y":
        y = np.full(40, 1.0)
    elif case == "n_equals_p":
        x, y, err, z = x[:3], y[:3], err[:3], z[:, :3]
    elif case == "zero_param_column":
        start = np.array([0.5, 0.0, 1.3])     # b = 0: the al column of the Jacobian vanishes (rank deficient)
    po, co, an = _run(emul, 0, x, y, err, start, z)
    assert po.shape == (8, 3)                 # it returned: every fit ended
    assert (co[:, :, 4] <= 2 * (200 + 150)).all()
    if case in ("constant_y", "n_equals_p", "zero_param_column"):
        assert np.isfinite(po).all()


_SPIN_CHILD = r"""
import signal, sys
sys.path.insert(0, %(tests)r)
import ctypes as C
import numpy as np
import oracle_lib as O
dp = C.POINTER(C.c_double)
L = O.lib()
L.fmc_oracle_tr_step.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, C.c_double, C.c_double, C.c_double, dp, dp,
                                 C.POINTER(C.c_int)]
eps = %(eps)r
# H = D^-1 (J^T J + S) D^-1 = diag(-1, 1), gradient g = J^T r = (eps^2, 0), radius 1
jac = np.array([eps, 0.0, 0.0, 0.0]); r = np.array([eps, 0.0]); d = np.ones(2); S = np.array([-1.0 - eps * eps, 0.0, 1.0])
step = np.zeros(2); v = np.zeros(9); ka = C.c_int(0)
D = lambda a: a.ctypes.data_as(dp)
print("start", flush=True)
signal.signal(signal.SIGALRM, signal.SIG_DFL)
signal.alarm(3)                                   # the default action kills the process inside the C call
L.fmc_oracle_tr_step(2, 2, 2, D(jac), D(r), D(d), D(S), 1.0, 0.0, 0.0, D(step), D(v), C.byref(ka))
signal.alarm(0)
print("returned", step.tolist(), ka.value, flush=True)
"""


def test_gqtstp_retry_loop_of_the_reference_has_no_bound_and_the_product_caps_it(emul):
    """The mechanism behind the one-in-3e7 GPU hang, reproduced on the CPU.

    gqtstp's retry after a failed Cholesky factorisation (toms573.f90:2828-2856, label 200) has no
    iteration limit.  With an indefinite H and a gradient that is negligible against its most
    negative eigenvalue -- the noise-level last iteration of a fit -- the upper bound
    uk = |g|/rad - emin rounds to -emin whenever the Gerschgorin estimate is tight, the safeguard
    sets alphak = uk * sqrt(lk/uk) = uk, H + uk I is singular to rounding, the factorisation fails
    with a zero pivot, lk := alphak, and the loop repeats with identical numbers for ever.  The
    restatement of the reference algorithm (the oracle) therefore never returns on
    H = diag(-1, 1), |g| = 1e-30; the product's solver leaves after 150 retries, flags the fit
    (anomaly bit 1) and returns a finite step."""
    import signal
    import subprocess
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    # sanity: with a gradient that matters the oracle's step returns and has the radius' length
    ok = subprocess.run([sys.executable, "-c", _SPIN_CHILD % dict(tests=here, eps=1e-3)], capture_output=True,
                        text=True, timeout=120)
    assert ok.returncode == 0 and "returned [-0.99999999" in ok.stdout, ok.stdout + ok.stderr
    spin = subprocess.run([sys.executable, "-c", _SPIN_CHILD % dict(tests=here, eps=1e-15)], capture_output=True,
                          text=True, timeout=120)
    assert "start" in spin.stdout and "returned" not in spin.stdout
    assert spin.returncode == -signal.SIGALRM                     # still inside the loop after 3 s

    emul.emul_gqtstp2.argtypes = [C.POINTER(C.c_double)] * 3 + [C.c_double, C.POINTER(C.c_double),
                                                                  C.POINTER(C.c_int), C.POINTER(C.c_int)]
    h, d = np.array([-1.0, 0.0, 1.0]), np.ones(2)
    for g0, capped in ((1e-30, True), (1e-16, True), (1e-6, False)):
        dig, step = np.array([g0, 0.0]), np.zeros(2)
        ka, an = C.c_int(0), C.c_int(0)
        emul.emul_gqtstp2(_d(h), _d(dig), _d(d), 1.0, _d(step), C.byref(ka), C.byref(an))
        assert np.isfinite(step).all() and np.linalg.norm(step) <= 1.1
        assert (an.value == 2 and ka.value == 150) if capped else (an.value == 0 and ka.value < 12)
        if not capped:
            assert np.allclose(step, [-1.0, 0.0])                 # along the negative-curvature direction
