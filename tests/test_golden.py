"""Golden fixtures produced by the reference's own Python implementation (mcfit.py; see
tests/golden/make_golden.py): the oracle -- and, on a GPU, the CUDA path -- must reproduce them.
mcfit.py fits with scipy's MINPACK Levenberg-Marquardt at its default tolerances (1.5e-8), a
different optimiser from NL2SOL, so the agreement bar is the optimisers' stopping accuracy."""
import json
import os
import re

import numpy as np
import pytest

import datasets

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold_stdout_numbers():
    t = open(os.path.join(GOLD, "mcfit_probe_stdout.txt")).read()
    opt = t.split("Optimised parameters")[1] if "Optimised parameters" in t else t
    vals = {}
    for name in ("a", "b", "alpha"):
        m = re.search(r"\b%s:\s*([-0-9.eE+]+)" % name, opt)
        vals[name] = float(m.group(1))
    chi = re.search(r"[Rr]educed chi\^2[^0-9\-]*([-0-9.eE+]+)", opt)
    return vals, (float(chi.group(1)) if chi else None), t


def test_oracle_reproduces_mcfit_optimised_parameters(oracle):
    vals, chi, text = _gold_stdout_numbers()
    model, x, y, err = datasets.dataset("c1")
    p, _ = oracle.fit_data(0, x, y, err, [0.0, 2.0, 1.0])
    assert np.allclose(p, [vals["a"], vals["b"], vals["alpha"]], rtol=5e-8)
    if chi is not None:
        r = oracle.madr(0, x, y, err, p)
        assert np.isclose((r * r).sum() / (len(x) - 3), chi, rtol=1e-10)


def test_oracle_matches_mcfit_curve_fit_on_seeded_resamples(oracle):
    """mcfit.py:96-99 on 40 seeded resamples of the probe data vs the oracle's fit_data."""
    g = json.load(open(os.path.join(GOLD, "mcfit_curve_fit.json")))
    model, x, y, err = datasets.dataset("c1")
    z = np.random.default_rng(g["seed"]).normal(size=(40, len(x)))
    start, _ = oracle.fit_data(0, x, y, err, [0.0, 2.0, 1.0])
    assert np.allclose(start, g["initparams"], rtol=5e-8)
    out = oracle.bootstrap(0, x, y, err, start, 40, z=z)
    ref = np.array(g["params"])
    sig = ref.std(0)
    assert np.allclose(out["params"], ref, rtol=1e-5)      # curve_fit (default tolerances) stops within ~2e-6
    assert (np.abs(out["params"] - ref) / sig).max() < 1e-4


def test_ranlux_fixture_matches_reference_comments(oracle):
    k = json.load(open(os.path.join(GOLD, "ranlux_kat.json")))
    oracle.rlux_reset_default()
    assert np.allclose(oracle.ranlux(100)[:5], k["default_seed_lux3"]["1-5"], atol=5.1e-9)
    assert np.allclose(oracle.ranlux(100)[:5], k["default_seed_lux3"]["101-105"], atol=5.1e-9)
    oracle.rluxgo(389, 1)
    assert np.allclose(oracle.ranlux(100)[:5], k["p389_seed1"]["1-5"], atol=5.1e-9)
    assert np.allclose(oracle.ranlux(100)[:5], k["p389_seed1"]["101-105"], atol=5.1e-9)


@pytest.mark.gpu
def test_cuda_path_matches_mcfit_curve_fit_on_seeded_resamples():
    import fitter_mc_b200 as fmc

    g = json.load(open(os.path.join(GOLD, "mcfit_curve_fit.json")))
    model, x, y, err = datasets.dataset("c1")
    z = np.random.default_rng(g["seed"]).normal(size=(40, len(x)))
    start, _ = fmc.fit_data(x, y, err, [0.0, 2.0, 1.0], 0)
    assert np.allclose(start, g["initparams"], rtol=5e-8)
    out = fmc.monte_carlo_fit_data(x, y, err, 0, nrand_points=40, params_init=start, initial_fit=False, z_inject=z,
                                   per_resample=True)
    ref = np.array(g["params"])
    assert np.allclose(out["params"], ref, rtol=1e-5)
    assert (np.abs(out["params"] - ref) / ref.std(0)).max() < 1e-4
