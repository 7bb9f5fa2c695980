"""BASELINE.json's single-GPU configuration at FULL size through the C ABI: properties that need no
oracle run of that size (additivity, reproducibility, consistency with the recorded 1e8 run)."""
import numpy as np
import pytest

import datasets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fmc():
    import fitter_mc_b200

    fitter_mc_b200.lib()
    return fitter_mc_b200


def _start(oracle, model, x, y, err):
    start, _ = oracle.fit_data(model, x, y, err, oracle.model_info(model)["params_init"])
    return start


def test_full_size_c2_step_properties(fmc, oracle):
    """BASELINE.json's single-GPU configuration at full size (C2: n=1000, 4 parameters, 1e6
    resamples): properties that do not need the oracle at that size -- additivity over a split of
    the id range, run-to-run reproducibility, every fit accounted for, and moments consistent with
    the recorded 1e8-resample run (profiles/r1_large_runs.json), of which ids [0, 1e6) are a subset."""
    model, x, y, err = datasets.dataset("c2")
    start = _start(oracle, model, x, y, err)
    n, cut = 1_000_000, 333_337
    kw = dict(params_init=start, initial_fit=False)
    whole = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=n, **kw)
    a = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=cut, first_counter=0, **kw)
    b = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=n - cut, first_counter=cut, **kw)
    again = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=n, **kw)
    assert int(whole["rc_count"].sum()) == n and np.array_equal(whole["rc_count"], a["rc_count"] + b["rc_count"])
    assert np.array_equal(whole["totals"], a["totals"] + b["totals"])
    for k in ("sum_p", "sum_p2", "sum_prop", "sum_prop2"):
        assert np.allclose(whole[k], a[k] + b[k], rtol=1e-12), k
        assert np.allclose(whole[k], again[k], rtol=1e-13), k              # only the order of atomics differs
    assert np.array_equal(whole["rc_count"], again["rc_count"])
    assert whole["rc_count"][3:7].sum() == n                               # every fit converged (codes 3..6)
    # the 1e8-resample run recorded in profiles/ (ids [0, 1e8), same seed, same data)
    mean8 = np.array([0.48876161034785753, 2.013872480958249, 1.287402992165191, 0.10219083461734821])
    sd8 = np.array([0.015156750989511062, 0.013717945889407598, 0.01926473137507048, 0.0036810568416719668])
    assert np.all(np.abs(whole["params_opt"] - mean8) < 5 * sd8 / np.sqrt(n))
    assert np.all(np.abs(whole["err_params_opt"] - sd8) < 5 * sd8 / np.sqrt(2 * n))
    assert abs(whole["props_opt"][0] - 2.5026340913061063) < 5 * 0.0091607551098173 / np.sqrt(n)
