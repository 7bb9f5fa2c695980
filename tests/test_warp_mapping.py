"""One fit per WARP (csrc/warp_kernel.cuh): the alternative mapping of BASELINE.json's north star
("one fit per thread or per warp, chosen by measurement"), opt-in through fmc_set_fit_mapping.

CPU part: the host image of the mapping (tests/host_emul) -- rows dealt to 32 lanes in groups of
four, per-lane partial sums, butterfly reduction in the order of the kernel's __shfl_xor rounds --
driving the product's solver: all lanes end every reduction with bitwise identical totals (the
property that keeps the warp's control flow uniform), and the fits match the oracle within the
parity gates and the one-fit-per-thread path to rounding.

GPU part: the CUDA kernel (two sweeps per pass, lane-major columns) through the C ABI against the
oracle, on every config -- 1024 fits of C4, the config this mapping is for."""
import ctypes as C
import os

import numpy as np
import pytest

import datasets
from test_analytic_jacobian import REL_MAX, REL_MEDIAN, SIGMA_MAX, _emul_run, emul  # noqa: F401

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return a.ctypes.data_as(_dp)


def _gates(name, got, want):
    rel = (np.abs(got - want) / np.abs(want)).max(1)
    sig = (np.abs(got - want) / want.std(0)).max()
    print("\n[warp %s] n=%d rel dev: median %.2e p99 %.2e max %.2e ; max %.2e sigma"
          % (name, len(got), np.median(rel), np.quantile(rel, 0.99), rel.max(), sig))
    assert np.median(rel) <= REL_MEDIAN and rel.max() <= REL_MAX and sig <= SIGMA_MAX


@pytest.mark.parametrize("jac", [0, 1])
@pytest.mark.parametrize("name,nres", [("c1", 300), ("c2", 60), ("c5mild", 60), ("c4", 6)])
def test_warp_mapping_host_image_matches_oracle_and_thread_mapping(emul, oracle, name, nres, jac):
    model, x, y, err = datasets.dataset(name)
    info = oracle.model_info(model)
    flags = oracle.LSVMIN_PER_FIT | (oracle.ANALYTIC_JAC if jac else 0)
    start, _ = oracle.fit_data(model, x, y, err, info["params_init"], flags=flags)
    z = np.random.default_rng(700 + model).normal(size=(nres, len(x)))
    want = oracle.bootstrap(model, x, y, err, start, nres, z=z, flags=flags, nthreads=8)
    emul.emul_bootstrap_warp.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_longlong, _dp, _dp, _ip]
    got = np.zeros((nres, info["nparam"]))
    cnt = np.zeros((nres, 2, 5), dtype=np.int32)
    disagreements = emul.emul_bootstrap_warp(model, jac, len(x), _d(x), _d(y), _d(err),
                                             _d(np.ascontiguousarray(start)), nres, _d(z), _d(got),
                                             cnt.ctypes.data_as(_ip))
    assert disagreements == 0              # every butterfly left the same bits in all 32 lanes
    _gates("%s jac=%d vs oracle" % (name, jac), got, want["params"])
    thread, _, tcnt, _ = _emul_run(emul, model, jac, x, y, err, start, z, info["nparam"], info["nprop"])
    rel = (np.abs(got - thread) / np.abs(thread)).max(1)
    print("[warp %s jac=%d] vs one fit per thread: median %.2e max %.2e ; same first-call history %.3f"
          % (name, jac, np.median(rel), rel.max(), (cnt[:, 0, :4] == tcnt[:, 0, :4]).all(1).mean()))
    assert np.median(rel) <= REL_MEDIAN and rel.max() <= REL_MAX
    assert (cnt[:, 0, :4] == tcnt[:, 0, :4]).all(1).mean() >= 0.8


def test_fit_mapping_argument_checks():
    import fitter_mc_b200 as fmc

    L = fmc.lib()
    assert L.fmc_set_fit_mapping(None, 5) == -4 and L.fmc_set_fit_mapping(None, 0) == 0
    # the warp-per-fit kernels are an optional part of the library (`make WARP=1`): accepted or refused, never ignored
    rc = L.fmc_set_fit_mapping(None, 1)
    assert rc in (0, -9)
    assert L.fmc_set_fit_mapping(None, 0) == 0


@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("name,nres", [("c1", 2000), ("c2p3", 400), ("c2", 400), ("c5mild", 300), ("c4", 1024)])
def test_cuda_warp_mapping_matches_oracle(oracle, name, nres):
    import fitter_mc_b200 as fmc

    if fmc.lib().fmc_set_fit_mapping(None, 1) == -9:
        pytest.skip("warp-per-fit kernels not in this build (lost the measurement, profiles/r2_mapping_*.json; make WARP=1)")
    fmc.lib().fmc_set_fit_mapping(None, 0)
    model, x, y, err = datasets.dataset(name)
    info = oracle.model_info(model)
    start, _ = oracle.fit_data(model, x, y, err, info["params_init"])
    z = np.random.default_rng(2700 + model).normal(size=(nres, len(x)))
    want = oracle.bootstrap(model, x, y, err, start, nres, z=z, nthreads=8)
    try:
        assert fmc.lib().fmc_set_fit_mapping(None, 1) == 0
        got = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=nres, params_init=start, initial_fit=False,
                                       z_inject=z, per_resample=True)
        gen = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=nres, params_init=start, initial_fit=False,
                                       per_resample=True)
    finally:
        fmc.lib().fmc_set_fit_mapping(None, 0)
    _gates("cuda/" + name, got["params"], want["params"])
    assert got["rc_count"].sum() == nres
    # the generator path deals the same deviates to the lanes as to the threads
    ref = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=nres, params_init=start, initial_fit=False,
                                   per_resample=True)
    rel = (np.abs(gen["params"] - ref["params"]) / np.abs(ref["params"])).max(1)
    assert np.median(rel) <= REL_MEDIAN and rel.max() <= REL_MAX
    assert np.allclose(gen["params_opt"], ref["params_opt"], rtol=1e-9)
