"""ctypes binding of libfmc_b200.so (include/fmc_b200.h) + the reference-facing host layer.

Reference interface mirrored here (bootstrap.f90):
  initialise_model      365-387   -> initialise_model(model)
  model_y               390-396   -> model_y(model, params, x)
  derived_properties    399-407   -> derived_properties(model, params)
  fit_data              215-226   -> fit_data(...)           (un-resampled fit, lines 288-291)
  monte_carlo_fit_data  253-362   -> monte_carlo_fit_data(...)  (lines 288-341; printing is the driver's job)
"""
import ctypes as C
import os

import numpy as np

DEFAULT_SEED = 31415  # utils.f90:7
_HERE = os.path.dirname(os.path.abspath(__file__))

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)


class FmcError(RuntimeError):
    def __init__(self, code, where):
        L = lib()
        msg = L.fmc_strerror(code).decode()
        detail = L.fmc_last_cuda_error().decode()
        super().__init__("%s: %s (%d)%s" % (where, msg, code, (" -- " + detail) if detail else ""))
        self.code = code


def lib_path():
    """The CUDA library: libfmc_b200.so next to this file, or the build named by FMC_B200_LIB."""
    return os.environ.get("FMC_B200_LIB") or os.path.join(_HERE, "libfmc_b200.so")


_lib = None


def lib():
    """Load the CUDA library.  Fails loudly if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError("libfmc_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C fitter_mc_b200 -j7`); fitter_mc_b200 has no CPU fallback")
    L = C.CDLL(path)
    L.fmc_strerror.restype = C.c_char_p
    L.fmc_strerror.argtypes = [C.c_int]
    L.fmc_last_cuda_error.restype = C.c_char_p
    L.fmc_version.restype = C.c_char_p
    L.fmc_model_count.restype = C.c_int
    L.fmc_initialise_model.argtypes = [C.c_int, _ip, _ip, _dp, C.c_void_p, C.c_void_p, C.c_char_p]
    L.fmc_model_y.restype = C.c_double
    L.fmc_model_y.argtypes = [C.c_int, _dp, C.c_double]
    L.fmc_derived_properties.argtypes = [C.c_int, _dp, _dp]
    L.fmc_bootstrap_run.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_longlong, C.c_ulonglong,
                                    C.c_longlong, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _lp, _lp]
    L.fmc_fit_once.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _ip]
    L.fmc_generate_offsets.argtypes = [C.c_ulonglong, C.c_longlong, C.c_longlong, C.c_int, _dp]
    L.fmc_debug_philox4x32.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.fmc_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.fmc_destroy.argtypes = [C.c_void_p]
    L.fmc_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.fmc_set_data.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp]
    L.fmc_set_accumulator_dev.argtypes = [C.c_void_p, C.c_void_p]
    L.fmc_accumulator_len.argtypes = [C.c_void_p]
    L.fmc_get_accumulator.argtypes = [C.c_void_p, _dp]
    L.fmc_launch.argtypes = [C.c_void_p, _dp, C.c_longlong, C.c_ulonglong, C.c_longlong, C.c_void_p, C.c_int]
    L.fmc_finish.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _lp, _lp]
    L.fmc_wait.argtypes = [C.c_void_p]
    L.fmc_last_anomalies.argtypes = [C.c_void_p, _lp, _ip, C.c_int]
    L.fmc_decode_accumulator.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _lp, _lp, _lp]
    L.fmc_last_kernel_ms.restype = C.c_float
    L.fmc_last_kernel_ms.argtypes = [C.c_void_p]
    L.fmc_last_launch_geometry.argtypes = [C.c_void_p, _ip]
    L.fmc_last_launch_count.restype = C.c_longlong
    L.fmc_last_launch_count.argtypes = [C.c_void_p]
    L.fmc_measure_dfma_peak.argtypes = [C.c_int, C.c_int, _dp, C.POINTER(C.c_float)]
    L.fmc_finalise_moments.argtypes = [C.c_int, C.c_longlong, _dp, _dp, _dp, _dp]
    L.fmc_bootstrap_run_binned.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_longlong, C.c_ulonglong,
                                           C.c_longlong, C.c_int, _dp, _dp, _dp, _dp, _lp, _lp, C.c_int, _dp, _dp,
                                           C.c_void_p]
    L.fmc_nccl_available.restype = C.c_int
    L.fmc_nccl_unique_id.argtypes = [C.c_char_p]
    L.fmc_nccl_init_rank.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
    L.fmc_allreduce.argtypes = [C.c_void_p]
    L.fmc_last_combine_method.restype = C.c_int
    L.fmc_set_fit_mapping.argtypes = [C.c_void_p, C.c_int]
    L.fmc_set_schedule.argtypes = [C.c_void_p, C.c_int]
    L.fmc_get_schedule.argtypes = [C.c_void_p]
    L.fmc_set_jacobian_mode.argtypes = [C.c_void_p, C.c_int]
    L.fmc_model_has_analytic_jacobian.argtypes = [C.c_int]
    L.fmc_set_histogram.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    L.fmc_histogram_len.restype = C.c_longlong
    L.fmc_histogram_len.argtypes = [C.c_void_p]
    L.fmc_set_histogram_dev.argtypes = [C.c_void_p, C.c_void_p]
    L.fmc_get_histogram.argtypes = [C.c_void_p, C.c_void_p]
    L.fmc_histogram_quantiles.restype = C.c_longlong
    L.fmc_histogram_quantiles.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, _dp, C.c_int, _dp]
    _lib = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _check(rc, where):
    if rc != 0:
        raise FmcError(rc, where)


def model_count():
    return lib().fmc_model_count()


def initialise_model(model=0):
    """bootstrap.f90:365-387: nparam, nprop, names and the initial parameter values."""
    nparam, nprop = C.c_int(), C.c_int()
    init = np.zeros(16)
    pn = C.create_string_buffer(16 * 3)
    qn = C.create_string_buffer(16 * 3)
    name = C.create_string_buffer(32)
    _check(lib().fmc_initialise_model(model, C.byref(nparam), C.byref(nprop), _d(init), pn, qn, name),
           "fmc_initialise_model")
    p, q = nparam.value, nprop.value
    return dict(name=name.value.decode(), nparam=p, nprop=q, params_init=init[:p].copy(),
                param_name=[pn.raw[3 * i:3 * i + 2].decode() for i in range(p)],
                prop_name=[qn.raw[3 * i:3 * i + 2].decode() for i in range(q)])


def model_y(model, params, x):
    """bootstrap.f90:390-396 (host image of the device callback)."""
    params = _f64(params)
    L = lib()
    xs = np.atleast_1d(np.asarray(x, dtype=np.float64))
    return np.array([L.fmc_model_y(model, _d(params), float(v)) for v in xs])


def derived_properties(model, params):
    """bootstrap.f90:399-407."""
    info = initialise_model(model)
    params = _f64(params)
    props = np.zeros(max(info["nprop"], 1))
    _check(lib().fmc_derived_properties(model, _d(params), _d(props)), "fmc_derived_properties")
    return props[:info["nprop"]]


def finalise_moments(nresample, s, s2):
    """bootstrap.f90:328-341."""
    s, s2 = _f64(s), _f64(s2)
    mean, err = np.empty(len(s)), np.empty(len(s))
    lib().fmc_finalise_moments(len(s), int(nresample), _d(s), _d(s2), _d(mean), _d(err))
    return mean, err


def fit_data(x, y, err_y, params, model=0):
    """offset_data(.FALSE.); fit_data(params) -- bootstrap.f90:288-291, on the GPU.
    Returns (params, [rc of the two nl2sno calls])."""
    x, y, err_y = map(_f64, (x, y, err_y))
    params = _f64(params).copy()
    rc2 = np.zeros(2, dtype=np.int32)
    _check(lib().fmc_fit_once(model, len(x), _d(x), _d(y), _d(err_y), _d(params), rc2.ctypes.data_as(_ip)),
           "fmc_fit_once")
    return params, rc2


def generate_offsets(seed, first_counter, nresample, ndata):
    z = np.empty((nresample, ndata))
    _check(lib().fmc_generate_offsets(seed, first_counter, nresample, ndata, _d(z)), "fmc_generate_offsets")
    return z


def debug_philox4x32(counters_keys):
    """Raw words of the device generator: counters_keys uint32 [n][6] (c0..c3, k0, k1) -> uint32 [n][4]."""
    ck = np.ascontiguousarray(counters_keys, dtype=np.uint32).reshape(-1, 6)
    out = np.zeros((len(ck), 4), dtype=np.uint32)
    _check(lib().fmc_debug_philox4x32(ck.ctypes.data_as(C.c_void_p), len(ck), out.ctypes.data_as(C.c_void_p)),
           "fmc_debug_philox4x32")
    return out


def decode_accumulator(model, params_start, acc):
    info = initialise_model(model)
    p, q = info["nparam"], info["nprop"]
    params_start, acc = _f64(params_start), _f64(acc)
    out = dict(sum_p=np.zeros(p), sum_p2=np.zeros(p), sum_prop=np.zeros(max(q, 1)), sum_prop2=np.zeros(max(q, 1)),
               rc_count=np.zeros(16, dtype=np.int64), totals=np.zeros(4, dtype=np.int64))
    nfit = C.c_longlong(0)
    _check(lib().fmc_decode_accumulator(model, _d(params_start), _d(acc), _d(out["sum_p"]), _d(out["sum_p2"]),
                                        _d(out["sum_prop"]), _d(out["sum_prop2"]), C.byref(nfit),
                                        out["rc_count"].ctypes.data_as(_lp), out["totals"].ctypes.data_as(_lp)),
           "fmc_decode_accumulator")
    out["nfit"] = nfit.value
    out["sum_prop"], out["sum_prop2"] = out["sum_prop"][:q], out["sum_prop2"][:q]
    return out


def monte_carlo_fit_data(x, y, err_y, model=0, nrand_points=100000, seed=DEFAULT_SEED, first_counter=0,
                         ngpus=1, params_init=None, z_inject=None, per_resample=False, initial_fit=True,
                         histogram_bins=0, histogram_range=None, pilot=4096):
    """monte_carlo_fit_data, bootstrap.f90:253-362, without the printing: initial fit on the raw
    data (288-295), the bootstrap loop (297-330, one C-ABI call), means and Bessel errors (328-341).

    histogram_bins > 0 is the large-N form of make_histogram (bootstrap.f90:301-325): every
    parameter and property is binned on the device; the result gains `hist_counts`
    [nparam + nprop][bins + 2], `hist_lo`, `hist_hi`.  Without histogram_range=(lo, hi) the range
    is mean -/+ 8 sigma of a pilot of `pilot` resamples (ids taken from the end of the id space
    so that they do not coincide with the run's own)."""
    info = initialise_model(model)
    p, q = info["nparam"], info["nprop"]
    x, y, err_y = map(_f64, (x, y, err_y))
    start = _f64(info["params_init"] if params_init is None else params_init).copy()
    rc_init = None
    if initial_fit:
        start, rc_init = fit_data(x, y, err_y, start, model)
    n = int(nrand_points)
    zi = None
    if z_inject is not None:
        zi = _f64(z_inject)
        assert zi.shape == (n, len(x))
    s = [np.zeros(p), np.zeros(p), np.zeros(max(q, 1)), np.zeros(max(q, 1))]
    pout = np.zeros((n, p)) if per_resample else None
    qout = np.zeros((n, max(q, 1))) if per_resample else None
    rcc = np.zeros(16, dtype=np.int64)
    tot = np.zeros(4, dtype=np.int64)
    hist = {}
    if histogram_bins:
        if zi is not None or per_resample:
            raise ValueError("binned runs keep no per-resample arrays; bin those with distributed.bin_values")
        nb = int(histogram_bins)
        if histogram_range is None:
            npilot = max(2, min(int(pilot), max(n, 2)))
            pp, pq = np.zeros((npilot, p)), np.zeros((npilot, max(q, 1)))
            dummy = [np.zeros(p), np.zeros(p), np.zeros(max(q, 1)), np.zeros(max(q, 1))]
            _check(lib().fmc_bootstrap_run(model, len(x), _d(x), _d(y), _d(err_y), _d(start), npilot, seed,
                                           (1 << 62) - npilot, 1, None, _d(dummy[0]), _d(dummy[1]), _d(dummy[2]),
                                           _d(dummy[3]), _d(pp), _d(pq), None, None), "fmc_bootstrap_run (pilot)")
            blo, bhi = histogram_range_from_pilot(np.column_stack([pp, pq[:, :q]]))
        else:
            blo, bhi = (_f64(v).copy() for v in histogram_range)
        counts = np.zeros((p + q, nb + 2), dtype=np.uint64)
        _check(lib().fmc_bootstrap_run_binned(model, len(x), _d(x), _d(y), _d(err_y), _d(start), n, seed,
                                              first_counter, ngpus, _d(s[0]), _d(s[1]), _d(s[2]), _d(s[3]),
                                              rcc.ctypes.data_as(_lp), tot.ctypes.data_as(_lp), nb, _d(blo), _d(bhi),
                                              counts.ctypes.data_as(C.c_void_p)), "fmc_bootstrap_run_binned")
        hist = dict(hist_counts=counts, hist_lo=blo, hist_hi=bhi)
    else:
        _check(lib().fmc_bootstrap_run(model, len(x), _d(x), _d(y), _d(err_y), _d(start), n, seed, first_counter,
                                       ngpus, _d(zi), _d(s[0]), _d(s[1]), _d(s[2]), _d(s[3]), _d(pout), _d(qout),
                                       rcc.ctypes.data_as(_lp), tot.ctypes.data_as(_lp)), "fmc_bootstrap_run")
    params_opt, err_params_opt = finalise_moments(n, s[0], s[1])
    props_opt, err_props_opt = finalise_moments(n, s[2][:q], s[3][:q])
    return dict(**hist, info=info, params_fit=start, props_fit=derived_properties(model, start), rc_init=rc_init,
                params_opt=params_opt, err_params_opt=err_params_opt, props_opt=props_opt,
                err_props_opt=err_props_opt, sum_p=s[0], sum_p2=s[1], sum_prop=s[2][:q], sum_prop2=s[3][:q],
                params=pout, props=None if qout is None else qout[:, :q], rc_count=rcc, totals=tot,
                nrand_points=n)


def histogram_quantiles(counts, lo, hi, probs):
    """Quantiles of one binned quantity (its nbins + 2 counts, fmc_b200.h).  Returns
    (quantiles, number of resamples outside [lo, hi))."""
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    probs = _f64(np.atleast_1d(probs))
    out = np.zeros(len(probs))
    outside = lib().fmc_histogram_quantiles(counts.ctypes.data_as(C.c_void_p), len(counts) - 2, float(lo), float(hi),
                                            _d(probs), len(probs), _d(out))
    if outside < 0:
        raise FmcError(-4, "fmc_histogram_quantiles")
    return out, int(outside)


def histogram_range_from_pilot(values, nsigma=8.0):
    """[lo, hi) per column from a pilot sample of per-resample results: mean -/+ nsigma standard
    deviations (a degenerate column gets a unit-width range round its value)."""
    v = np.asarray(values, dtype=np.float64)
    v = v[np.all(np.isfinite(v), axis=1)]
    m, sd = v.mean(axis=0), v.std(axis=0)
    sd = np.where(sd > 0, sd, 0.5 / nsigma)
    return m - nsigma * sd, m + nsigma * sd


def nccl_available():
    """0 if libnccl.so.2 cannot be loaded, else NCCL's version code."""
    return int(lib().fmc_nccl_available())


def nccl_unique_id():
    """128 bytes for rank 0 to broadcast to the other ranks (Context.nccl_init_rank)."""
    buf = C.create_string_buffer(128)
    _check(lib().fmc_nccl_unique_id(buf), "fmc_nccl_unique_id")
    return buf.raw


def last_combine_method():
    """How the last one-call multi-GPU run combined its GPUs: "none", "nccl" or "host"."""
    return {0: "none", 1: "nccl", 2: "host"}[int(lib().fmc_last_combine_method())]


def measure_dfma_peak(device=0, iters=20000):
    """Measured FP64 FMA peak of the device in TFLOP/s (register-resident DFMA chains)."""
    t = C.c_double(0.0)
    ms = C.c_float(0.0)
    _check(lib().fmc_measure_dfma_peak(device, iters, C.byref(t), C.byref(ms)), "fmc_measure_dfma_peak")
    return t.value, ms.value


class Context:
    """One GPU: resident data, a stream, asynchronous launches (fmc_create ... fmc_finish)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(lib().fmc_create(C.byref(self._h), device), "fmc_create")
        self.model = None
        self.device = device

    def close(self):
        if self._h:
            lib().fmc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        _check(lib().fmc_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), "fmc_set_stream")

    def set_data(self, model, x, y, err_y):
        x, y, err_y = map(_f64, (x, y, err_y))
        _check(lib().fmc_set_data(self._h, model, len(x), _d(x), _d(y), _d(err_y)), "fmc_set_data")
        self.model = model
        self.info = initialise_model(model)
        self.ndata = len(x)

    def accumulator_len(self):
        return lib().fmc_accumulator_len(self._h)

    def read_accumulator(self):
        """The accumulator block on the host, after the work queued so far (launch, allreduce)."""
        acc = np.zeros(self.accumulator_len())
        _check(lib().fmc_get_accumulator(self._h, _d(acc)), "fmc_get_accumulator")
        return acc

    def set_accumulator_dev(self, ptr):
        _check(lib().fmc_set_accumulator_dev(self._h, C.c_void_p(ptr)), "fmc_set_accumulator_dev")

    def launch(self, params_start, nresample, seed=DEFAULT_SEED, first_counter=0, z_inject_dev=None, keep=False):
        self._start = _f64(params_start).copy()
        self._nres = int(nresample)
        self._keep = keep
        _check(lib().fmc_launch(self._h, _d(self._start), self._nres, seed, first_counter,
                                C.c_void_p(z_inject_dev) if z_inject_dev else None, 1 if keep else 0), "fmc_launch")

    def finish(self):
        p, q, n = self.info["nparam"], self.info["nprop"], self._nres
        s = [np.zeros(p), np.zeros(p), np.zeros(max(q, 1)), np.zeros(max(q, 1))]
        pout = np.zeros((n, p)) if self._keep else None
        qout = np.zeros((n, max(q, 1))) if self._keep else None
        cout = np.zeros((n, 2, 5), dtype=np.int32) if self._keep else None
        rcc = np.zeros(16, dtype=np.int64)
        tot = np.zeros(4, dtype=np.int64)
        _check(lib().fmc_finish(self._h, _d(s[0]), _d(s[1]), _d(s[2]), _d(s[3]), _d(pout), _d(qout),
                                None if cout is None else cout.ctypes.data_as(_ip),
                                rcc.ctypes.data_as(_lp), tot.ctypes.data_as(_lp)), "fmc_finish")
        return dict(sum_p=s[0], sum_p2=s[1], sum_prop=s[2][:q], sum_prop2=s[3][:q], params=pout,
                    props=None if qout is None else qout[:, :q], counters=cout, rc_count=rcc, totals=tot)

    def wait(self):
        _check(lib().fmc_wait(self._h), "fmc_wait")

    def nccl_init_rank(self, unique_id, nranks, rank):
        """Bind this GPU as rank `rank` of `nranks` (unique_id: nccl_unique_id() of rank 0)."""
        assert len(unique_id) == 128
        _check(lib().fmc_nccl_init_rank(self._h, unique_id, int(nranks), int(rank)), "fmc_nccl_init_rank")

    def allreduce(self):
        """The single collective of the path: one ncclAllReduce of the accumulator block (and of the
        binned counts) over the bound ranks, enqueued on the context's stream after launch()."""
        _check(lib().fmc_allreduce(self._h), "fmc_allreduce")

    def set_jacobian_mode(self, mode):
        """"forward-difference" (nl2sno, the reference's path; default) or "analytic" (nl2sol + madj,
        README:35-42): exact Jacobian rows from the plug-in's model_grad or by dual numbers."""
        code = {"forward-difference": 0, "fd": 0, 0: 0, "analytic": 1, 1: 1}[mode]
        _check(lib().fmc_set_jacobian_mode(self._h, code), "fmc_set_jacobian_mode")

    def set_fit_mapping(self, mapping):
        """"thread" (default: one resample per thread) or "warp" (one resample per warp; experimental,
        meant for large ndata x nparam -- csrc/warp_kernel.cuh)."""
        code = {"thread": 0, 0: 0, "warp": 1, 1: 1}[mapping]
        _check(lib().fmc_set_fit_mapping(self._h, code), "fmc_set_fit_mapping")

    def set_schedule(self, schedule):
        """"refill" (one persistent kernel per nl2sno call, threads take the next fit from a counter) or
        "rounds" (solver steps and passes in separate kernels alternating inside a CUDA graph WHILE node,
        fits compacted between rounds -- csrc/round_kernels.cuh).  Same per-fit arithmetic."""
        code = {"refill": 0, 0: 0, "rounds": 1, 1: 1, "auto": 2, 2: 2}[schedule]
        _check(lib().fmc_set_schedule(self._h, code), "fmc_set_schedule")

    def schedule(self):
        return {0: "refill", 1: "rounds", 2: "auto"}[lib().fmc_get_schedule(self._h)]

    def set_histogram(self, nbins, lo=None, hi=None):
        """Bin every fitted parameter and property on the device (replaces histogram_<prop>.dat,
        bootstrap.f90:319-325).  lo/hi: [nparam + nprop], parameters first.  nbins=0: off."""
        if nbins:
            lo, hi = _f64(lo), _f64(hi)
            k = self.info["nparam"] + self.info["nprop"]
            if lo.shape != (k,) or hi.shape != (k,):
                raise ValueError("lo and hi need nparam + nprop = %d entries" % k)
        _check(lib().fmc_set_histogram(self._h, int(nbins), _d(lo) if nbins else None, _d(hi) if nbins else None),
               "fmc_set_histogram")
        self._hist = (int(nbins), lo, hi) if nbins else None

    def histogram_len(self):
        return int(lib().fmc_histogram_len(self._h))

    def set_histogram_dev(self, ptr):
        _check(lib().fmc_set_histogram_dev(self._h, C.c_void_p(ptr)), "fmc_set_histogram_dev")

    def histogram(self):
        """Counts of the last launch on this GPU, uint64 [nparam + nprop][nbins + 2]."""
        nbins = self._hist[0]
        out = np.zeros((self.histogram_len() // (nbins + 2), nbins + 2), dtype=np.uint64)
        _check(lib().fmc_get_histogram(self._h, out.ctypes.data_as(C.c_void_p)), "fmc_get_histogram")
        return out

    def anomalies(self):
        """(ids, flags) of fits of the last launch that hit a solver safety cap (normally none)."""
        ids = np.zeros(64, dtype=np.int64)
        flags = np.zeros(64, dtype=np.int32)
        n = lib().fmc_last_anomalies(self._h, ids.ctypes.data_as(_lp), flags.ctypes.data_as(_ip), 64)
        if n < 0:
            raise FmcError(n, "fmc_last_anomalies")
        return ids[:min(n, 64)], flags[:min(n, 64)], n

    def kernel_ms(self):
        return float(lib().fmc_last_kernel_ms(self._h))

    def launch_count(self):
        """Kernels of this library that the last launch() enqueued."""
        return int(lib().fmc_last_launch_count(self._h))

    def geometry(self):
        g = np.zeros(4, dtype=np.int32)
        lib().fmc_last_launch_geometry(self._h, g.ctypes.data_as(_ip))
        return dict(blocks=int(g[0]), threads=int(g[1]), blocks_per_sm=int(g[2]), sms=int(g[3]))
