"""The reference's own reproducibility floor, measured: the SAME algorithm (the oracle's
restatement of nl2sno) compiled (a) strictly (-O2 -ffp-contract=off) and (b) with the
reference's default flags (-Ofast -fpeel-loops -march=native, makedefs.mk:16-18) is run on
identical injected resamples.  The two builds differ per fit by ~1e-9 (median) to ~7e-8 (max)
relative: the forward-difference Jacobian amplifies rounding to ~1e-8, and the last,
noise-level step of the second nl2sno call is accepted or not depending on the sign of a
rounding-level change of f.  BASELINE.json's 1e-9 per-fit tolerance is therefore below what the
reference can reproduce of itself; tests/test_gpu_parity.py asserts bounds of this size."""
import os
import subprocess

import numpy as np
import pytest

import datasets


@pytest.fixture(scope="module")
def ofast_lib(tmp_path_factory, oracle):
    out = str(tmp_path_factory.mktemp("ofast") / "libfmc_oracle_ofast.so")
    src = [os.path.join(oracle.ORACLE_DIR, f) for f in ("nl2sol_port.cpp", "ranlux_port.cpp", "bootstrap_port.cpp",
                                                       "test_hooks.cpp")]
    subprocess.check_call(["/usr/bin/g++", "-Ofast", "-fpeel-loops", "-march=native", "-mtune=native", "-fPIC",
                           "-fopenmp", "-std=c++17", "-w", "-shared", "-o", out] + src)
    return out


def _run(oracle, path, model, x, y, err, start, n, z):
    import ctypes

    saved = oracle._lib
    try:
        oracle._lib = None
        real_build, real_path = oracle.build, oracle.LIB_PATH
        oracle.build, oracle.LIB_PATH = (lambda: None), path
        return oracle.bootstrap(model, x, y, err, start, n, z=z, nthreads=4)
    finally:
        oracle.build, oracle.LIB_PATH = real_build, real_path
        oracle._lib = saved


@pytest.mark.parametrize("name,nres", [("c1", 1500), ("c2", 300)])
def test_reference_algorithm_is_not_reproducible_to_1e9(oracle, ofast_lib, name, nres):
    model, x, y, err = datasets.dataset(name)
    start, _ = oracle.fit_data(model, x, y, err, oracle.model_info(model)["params_init"])
    z = np.random.default_rng(1).normal(size=(nres, len(x)))
    a = oracle.bootstrap(model, x, y, err, start, nres, z=z, nthreads=4)
    b = _run(oracle, ofast_lib, model, x, y, err, start, nres, z)
    rel = (np.abs(a["params"] - b["params"]) / np.abs(a["params"])).max(1)
    sig = (np.abs(a["params"] - b["params"]) / a["params"].std(0)).max()
    print("\n[floor %s] strict vs -Ofast build of the same algorithm: median %.2e p99 %.2e max %.2e ; %.2e sigma"
          % (name, np.median(rel), np.quantile(rel, 0.99), rel.max(), sig))
    assert rel.max() > 1e-9          # the 1e-9 bar is not met by the reference against itself
    assert np.median(rel) < 2e-8 and rel.max() < 1e-6 and sig < 1e-4   # ... but the floor is tiny
