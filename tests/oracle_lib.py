"""ctypes binding of the CPU parity oracle (oracle/libfmc_oracle.so).

TEST INFRASTRUCTURE: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs only.  The product package never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libfmc_oracle.so")

COVARIANCE, COPY_JAC, LSVMIN_PER_FIT, HISTOGRAM, ANALYTIC_JAC, SAFETY_CAP = 1, 2, 4, 8, 16, 32
RNG_RANLUX, RNG_INJECT = 0, 1
NCOUNTERS = 6

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)


def build():
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".hpp", ".h"))]
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return
    subprocess.check_call(["make", "-C", ORACLE_DIR, "libfmc_oracle.so"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.fmc_oracle_rluxgo.argtypes = [C.c_int] * 4
        L.fmc_oracle_ranlux.argtypes = [_dp, C.c_int]
        L.fmc_oracle_rluxat_kount.restype = C.c_longlong
        L.fmc_oracle_rang.argtypes = [C.c_int, _dp]
        L.fmc_oracle_model_y.restype = C.c_double
        L.fmc_oracle_model_y.argtypes = [C.c_int, _dp, C.c_double]
        L.fmc_oracle_derived_properties.argtypes = [C.c_int, _dp, _dp]
        L.fmc_oracle_initialise_model.argtypes = [C.c_int, _ip, _ip, _dp, C.c_void_p, C.c_void_p, C.c_char_p]
        L.fmc_oracle_fit_data.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_int, _ip]
        L.fmc_oracle_madr.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.fmc_oracle_madj.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        L.fmc_oracle_bootstrap.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_longlong, C.c_int, _dp,
                                           C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _lp, _lp]
        L.fmc_oracle_finalise_moments.argtypes = [C.c_int, C.c_longlong, _dp, _dp, _dp, _dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def rluxgo(lux, seed, k1=0, k2=0):
    lib().fmc_oracle_rluxgo(lux, seed, k1, k2)


def rlux_reset_default():
    lib().fmc_oracle_rlux_reset_default()


def ranlux(n):
    out = np.empty(n)
    lib().fmc_oracle_ranlux(_d(out), n)
    return out


def initialise_rng():
    lib().fmc_oracle_initialise_rng()


def rang(n):
    out = np.empty(n)
    lib().fmc_oracle_rang(n, _d(out))
    return out


def model_info(model):
    nparam, nprop = C.c_int(), C.c_int()
    init = np.zeros(16)
    pn = C.create_string_buffer(16 * 3)
    qn = C.create_string_buffer(16 * 3)
    name = C.create_string_buffer(32)
    rc = lib().fmc_oracle_initialise_model(model, C.byref(nparam), C.byref(nprop), _d(init), pn, qn, name)
    if rc != 0:
        raise ValueError("bad model id %d" % model)
    p, q = nparam.value, nprop.value
    return dict(name=name.value.decode(), nparam=p, nprop=q, params_init=init[:p].copy(),
                param_name=[pn.raw[3 * i:3 * i + 2].decode() for i in range(p)],
                prop_name=[qn.raw[3 * i:3 * i + 2].decode() for i in range(q)])


def model_y(model, params, x):
    params = _f64(params)
    return np.array([lib().fmc_oracle_model_y(model, _d(params), float(xi)) for xi in np.atleast_1d(x)])


def derived_properties(model, params):
    params = _f64(params)
    out = np.zeros(max(model_info(model)["nprop"], 1))
    lib().fmc_oracle_derived_properties(model, _d(params), _d(out))
    return out[:model_info(model)["nprop"]]


def madr(model, x, y_fit, err, params):
    x, y_fit, err, params = map(_f64, (x, y_fit, err, params))
    r = np.empty(len(x))
    lib().fmc_oracle_madr(model, len(x), _d(x), _d(y_fit), _d(err), _d(params), _d(r))
    return r


def madj(model, x, err, params):
    """Analytic Jacobian of madr, [n][p] (hand-derived formulas, the user's `madj` of README:35-42)."""
    x, err, params = map(_f64, (x, err, params))
    j = np.empty((len(params), len(x)))
    lib().fmc_oracle_madj(model, len(x), _d(x), _d(err), _d(params), _d(j))
    return j.T.copy()


def fit_data(model, x, y_fit, err, params, flags=LSVMIN_PER_FIT):
    """bootstrap.f90:215-226.  Returns (params, counters[2][6])."""
    x, y_fit, err = map(_f64, (x, y_fit, err))
    params = _f64(params).copy()
    counters = np.zeros((2, NCOUNTERS), dtype=np.int32)
    rc = lib().fmc_oracle_fit_data(model, len(x), _d(x), _d(y_fit), _d(err), _d(params), flags,
                                   counters.ctypes.data_as(_ip))
    if rc != 0:
        raise RuntimeError("fmc_oracle_fit_data rc=%d" % rc)
    return params, counters


def bootstrap(model, x, y, err, params_start, nresample, z=None, flags=LSVMIN_PER_FIT, nthreads=1,
              per_resample=True):
    """bootstrap.f90:297-330.  z=None -> the serial RANLUX+polar stream (call initialise_rng() first)."""
    info = model_info(model)
    p, q = info["nparam"], info["nprop"]
    x, y, err, params_start = map(_f64, (x, y, err, params_start))
    n = len(x)
    if z is not None:
        z = _f64(z)
        assert z.shape == (nresample, n)
    sums = [np.zeros(p), np.zeros(p), np.zeros(max(q, 1)), np.zeros(max(q, 1))]
    pout = np.zeros((nresample, p)) if per_resample else None
    qout = np.zeros((nresample, max(q, 1))) if per_resample else None
    cout = np.zeros((nresample, 2, NCOUNTERS), dtype=np.int32) if per_resample else None
    rcc = np.zeros(16, dtype=np.int64)
    nev = C.c_longlong(0)
    rc = lib().fmc_oracle_bootstrap(
        model, n, _d(x), _d(y), _d(err), _d(params_start), nresample,
        RNG_INJECT if z is not None else RNG_RANLUX, _d(z) if z is not None else None, flags, nthreads,
        _d(sums[0]), _d(sums[1]), _d(sums[2]), _d(sums[3]),
        _d(pout) if per_resample else None, _d(qout) if per_resample else None,
        cout.ctypes.data_as(_ip) if per_resample else None, rcc.ctypes.data_as(_lp), C.byref(nev))
    if rc != 0:
        raise RuntimeError("fmc_oracle_bootstrap rc=%d" % rc)
    return dict(sum_p=sums[0], sum_p2=sums[1], sum_prop=sums[2][:q], sum_prop2=sums[3][:q],
                params=pout, props=None if qout is None else qout[:, :q], counters=cout, rc_count=rcc,
                neval=nev.value)


def finalise_moments(nresample, s, s2):
    s, s2 = _f64(s), _f64(s2)
    mean, err = np.empty(len(s)), np.empty(len(s))
    lib().fmc_oracle_finalise_moments(len(s), nresample, _d(s), _d(s2), _d(mean), _d(err))
    return mean, err
