"""Experimental shared-first-trial path (csrc/trial1_kernel.cuh, off unless
FMC_SHARED_FIRST_TRIAL=1): in the first solver call every fit's first trial pass has
x0 = params_start, so J(x0)^T r(x) can come from the Jacobian rows jac0_kernel evaluates once.

CPU part: the host image of that pass inside the product's solver -- same fits as the normal
flow up to rounding, and within the parity gates of the oracle.  GPU part (gated like the
warp-mapping tests: the kernels have not run on a GPU yet): the CUDA path with the switch on
equals the CUDA path with it off up to rounding."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import datasets
from test_analytic_jacobian import REL_MAX, REL_MEDIAN, SIGMA_MAX, _emul_run, emul  # noqa: F401

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _d(a):
    return a.ctypes.data_as(_dp)


@pytest.mark.parametrize("name,nres", [("c1", 400), ("c2", 80), ("c5mild", 80), ("c4", 6)])
def test_shared_first_trial_host_image(emul, oracle, name, nres):
    model, x, y, err = datasets.dataset(name)
    info = oracle.model_info(model)
    start, _ = oracle.fit_data(model, x, y, err, info["params_init"])
    z = np.random.default_rng(3100 + model).normal(size=(nres, len(x)))
    want = oracle.bootstrap(model, x, y, err, start, nres, z=z, nthreads=8)
    emul.emul_bootstrap_trial1.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_longlong, _dp, _dp, _ip]
    emul.emul_trial1_lky_diff.restype = C.c_double
    emul.emul_trial1_lky_diff.argtypes = [C.c_int]
    emul.emul_trial1_lky_diff(1)
    got = np.zeros((nres, info["nparam"]))
    cnt = np.zeros((nres, 2, 5), dtype=np.int32)
    assert emul.emul_bootstrap_trial1(model, len(x), _d(x), _d(y), _d(err), _d(np.ascontiguousarray(start)), nres,
                                      _d(z), _d(got), cnt.ctypes.data_as(_ip)) == 0
    rel = (np.abs(got - want["params"]) / np.abs(want["params"])).max(1)
    sig = (np.abs(got - want["params"]) / want["params"].std(0)).max()
    assert np.median(rel) <= REL_MEDIAN and rel.max() <= REL_MAX and sig <= SIGMA_MAX
    normal, _, ncnt, _ = _emul_run(emul, model, 0, x, y, err, start, z, info["nparam"], info["nprop"])
    reln = (np.abs(got - normal) / np.abs(normal)).max(1)
    same = (cnt[:, 0, :4] == ncnt[:, 0, :4]).all(1).mean()
    print("\n[shared first trial %s] vs oracle: median %.2e max %.2e ; vs normal flow: median %.2e max %.2e ; "
          "same first-call history %.3f" % (name, np.median(rel), rel.max(), np.median(reln), reln.max(), same))
    assert np.median(reln) <= REL_MEDIAN and reln.max() <= REL_MAX and same >= 0.8
    # In these small-residual fits the secant matrix S never decides a step, so the fits above would
    # agree even with a wrong lky; compare the vector itself with the normal flow's J(x0)^T r(x).
    lky_diff = emul.emul_trial1_lky_diff(1)
    print("[shared first trial %s] worst relative difference of lky: %.2e" % (name, lky_diff))
    assert 0.0 < lky_diff <= 1e-11


_CHILD = r"""
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import numpy as np
import fitter_mc_b200 as fmc, datasets, oracle_lib as oracle
model, x, y, err = datasets.dataset(%(name)r)
start, _ = oracle.fit_data(model, x, y, err, fmc.initialise_model(model)["params_init"])
out = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=%(nres)d, params_init=start, initial_fit=False,
                               per_resample=True)
np.save(%(path)r, out["params"])
"""


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("FMC_TEST_TRIAL1_PATH") != "1",
                    reason="shared-first-trial kernels not yet run on a GPU (set FMC_TEST_TRIAL1_PATH=1)")
@pytest.mark.timeout(300)
@pytest.mark.parametrize("name,nres", [("c1", 20000), ("c2", 5000), ("c5mild", 2000)])
def test_cuda_shared_first_trial_equals_default_path(tmp_path, name, nres):
    res = {}
    for flag in ("0", "1"):                       # the switch is read when a context is created: fresh processes
        path = str(tmp_path / ("p%s.npy" % flag))
        env = dict(os.environ, FMC_SHARED_FIRST_TRIAL=flag)
        p = subprocess.run([sys.executable, "-c", _CHILD % dict(root=ROOT, name=name, nres=nres, path=path)],
                           capture_output=True, text=True, timeout=240, env=env)
        assert p.returncode == 0, p.stderr[-2000:]
        res[flag] = np.load(path)
    rel = (np.abs(res["1"] - res["0"]) / np.abs(res["0"])).max(1)
    print("\n[cuda shared first trial %s] vs default: median %.2e max %.2e" % (name, np.median(rel), rel.max()))
    assert np.median(rel) <= REL_MEDIAN and rel.max() <= REL_MAX
    assert not np.array_equal(res["1"], res["0"])   # it really took the other route
