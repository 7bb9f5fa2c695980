"""BASELINE.json configs[4] / BASELINE.md C5: "stiff exponential-plus-power-law model with derived
properties and HIGH NON-CONVERGENCE RATE (divergence/compaction stress)", err tuned "until >= 10 %
of fits end with NL2SOL codes 7-10 (measured by the oracle)".

What the reference does with such fits: nothing special -- fit_data never inspects nl2sno's
return code (bootstrap.f90:215-226; codes at toms573.f90:265-307, limits mxiter/mxfcal at
1010-1013 and 1053-1054), the parameters go into the sums like any other.

Measured with the oracle on tests/datasets.py:c5 (err = 13|y| + 1, 1000 injected resamples,
seed 1):  second-call return codes 4: 766, 7 (singular convergence): 223, 8 (false convergence): 4,
9 (evaluation limit): 3, 10 (iteration limit): 3  ->  23.3 % codes 7-10; 5 ... 500 passes per fit
(mean 77, median 28, 90 % <= 272) against 6.6 on C2.

Parity criterion.  In this regime the exponential term is not identifiable (a runs to -1e16 along a
flat valley), so WHERE a fit stops in the valley is decided at rounding level: the reference's own
-O2 and -Ofast builds do not agree on parameters here, nor can any other implementation.  What every
correct implementation of the algorithm does agree on is the chi-squared it reaches: the tests
compare chi^2(params) of the CUDA path (and of the host image of its solver) with the oracle's on
identical injected resamples -- tightly for fits both codes call converged (codes 3-6), loosely
for those that end non-converged -- and the return-code mix.  Per-fit PARAMETER parity of this
model is asserted on the untuned recipe ("c5mild": tests/test_gpu_parity.py,
tests/test_host_emul.py, tests/test_analytic_jacobian.py).
"""
import ctypes as C

import numpy as np
import pytest

import datasets
from test_analytic_jacobian import _emul_run, emul  # noqa: F401

NONCONV_MIN = 0.10


def _chi2(x, y, err, z, p):
    with np.errstate(all="ignore"):
        yf = y[None, :] + z * err[None, :]
        m = p[:, 0:1] * np.exp(-p[:, 1:2] * x[None, :]) + p[:, 2:3] * np.exp(-p[:, 3:4] * np.log(x)[None, :])
        return (((yf - m) / err[None, :]) ** 2).sum(1)


def _setup(oracle, nres, seed=1):
    model, x, y, err = datasets.dataset("c5")
    start = datasets.C5_FITTED.copy()     # committed: the initial fit itself is chaotic (tests/datasets.py)
    z = np.random.default_rng(seed).normal(size=(nres, len(x)))
    return model, x, y, err, start, z


def _oracle_flags(oracle):
    # the checker must not hang the suite on this data: bound the reference's unbounded gqtstp retry loop
    return oracle.LSVMIN_PER_FIT | oracle.SAFETY_CAP


def _chi2_gates(name, x, y, err, z, got, want, rc_got):
    """chi^2 reached by `got` vs the oracle's `want` on the same resamples."""
    fg, fo = _chi2(x, y, err, z, got), _chi2(x, y, err, z, want)
    finite = np.isfinite(fg) & np.isfinite(fo)
    d = np.abs(fg - fo)[finite] / fo[finite]
    conv = (rc_got <= 6)[finite]
    print("\n[c5 stress %s] n=%d  non-finite chi2 on either side: %d" % (name, len(got), (~finite).sum()))
    for lab, sel in (("both converged (codes 3-6)", conv), ("non-converged (codes 7-10)", ~conv)):
        dd = d[sel]
        print("   %-28s %4d fits  chi2 rel diff median %.2e  p90 %.2e  p99 %.2e  max %.2e"
              % (lab, sel.sum(), np.median(dd), np.quantile(dd, 0.9), np.quantile(dd, 0.99), dd.max()))
    assert (~finite).mean() <= 0.01
    assert np.median(d[conv]) <= 1e-12 and np.quantile(d[conv], 0.99) <= 1e-8 and (d[conv] > 1e-6).mean() <= 0.01
    assert np.median(d[~conv]) <= 1e-5 and np.quantile(d[~conv], 0.99) <= 1e-2
    assert d.max() <= 0.05


def test_c5_recipe_gives_a_high_non_convergence_rate(oracle):
    model, x, y, err, start, z = _setup(oracle, 1000)
    res = oracle.bootstrap(model, x, y, err, start, len(z), z=z, flags=_oracle_flags(oracle), nthreads=8)
    rc = res["rc_count"]
    frac = rc[7:11].sum() / len(z)
    print("\n[c5 stress] oracle return codes 3..10: %s  -> %.1f %% codes 7-10" % (rc[3:11].tolist(), 100 * frac))
    assert rc.sum() == len(z)
    assert frac >= NONCONV_MIN
    # the committed start is what the oracle's initial fit gives in the build container
    refit, _ = oracle.fit_data(model, x, y, err, oracle.model_info(model)["params_init"], flags=_oracle_flags(oracle))
    print("[c5 stress] initial fit here: %s ; committed: %s" % (refit.tolist(), start.tolist()))
    # the untuned recipe has none (which is why the noise was tuned up)
    model, x, y, err = datasets.dataset("c5mild")
    start, _ = oracle.fit_data(model, x, y, err, oracle.model_info(model)["params_init"])
    mild = oracle.bootstrap(model, x, y, err, start, 200, z=z[:200], nthreads=8)
    assert mild["rc_count"][7:11].sum() == 0


def test_host_image_of_the_solver_reaches_the_oracles_chi2_on_c5(emul, oracle):
    model, x, y, err, start, z = _setup(oracle, 240)
    want = oracle.bootstrap(model, x, y, err, start, len(z), z=z, flags=_oracle_flags(oracle), nthreads=8)
    got, _, cnt, anomalies = _emul_run(emul, model, 0, x, y, err, start, z, 4, 2)
    rc = cnt[:, 1, 0]
    _chi2_gates("host image", x, y, err, z, got, want["params"], rc)
    assert (anomalies != 0).sum() == 0                      # no safety cap needed
    assert (rc >= 7).mean() >= NONCONV_MIN
    assert abs((rc >= 7).mean() - want["rc_count"][7:11].sum() / len(z)) <= 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("mapping,schedule", [("thread", "refill"), ("thread", "rounds"), ("warp", "refill")])
def test_cuda_path_reaches_the_oracles_chi2_on_c5(oracle, mapping, schedule):
    import fitter_mc_b200 as fmc

    if mapping == "warp" and fmc.lib().fmc_set_fit_mapping(None, 1) == -9:
        pytest.skip("warp-per-fit kernels not in this build (make WARP=1)")
    fmc.lib().fmc_set_fit_mapping(None, 0)
    model, x, y, err, start, z = _setup(oracle, 1000)
    want = oracle.bootstrap(model, x, y, err, start, len(z), z=z, flags=_oracle_flags(oracle), nthreads=8)
    zdev = None
    ctx = fmc.Context(0)
    try:
        import torch

        ctx.set_data(model, x, y, err)
        ctx.set_fit_mapping(mapping)
        ctx.set_schedule(schedule)
        zdev = torch.from_numpy(z).cuda()
        torch.cuda.synchronize()
        ctx.launch(start, len(z), z_inject_dev=zdev.data_ptr(), keep=True)
        got = ctx.finish()
        ids, flags, n_anom = ctx.anomalies()
    finally:
        ctx.close()
    rc = got["counters"][:, 1, 0]
    passes = got["counters"][:, :, 4].sum(1)
    mapping = mapping + "/" + schedule
    print("\n[c5 stress cuda/%s] return codes 3..10: gpu %s oracle %s ; passes per fit mean %.1f median %d p90 %d "
          "p99 %d max %d ; safety-cap hits %d"
          % (mapping, np.bincount(rc, minlength=11)[3:11].tolist(), want["rc_count"][3:11].tolist(), passes.mean(),
             np.median(passes), np.quantile(passes, 0.9), np.quantile(passes, 0.99), passes.max(), n_anom))
    _chi2_gates("cuda/" + mapping, x, y, err, z, got["params"], want["params"], rc)
    assert got["rc_count"].sum() == len(z)
    assert (rc >= 7).mean() >= NONCONV_MIN
    assert abs((rc >= 7).mean() - want["rc_count"][7:11].sum() / len(z)) <= 0.04
    assert n_anom <= 2
    # non-converged fits are averaged like any other (bootstrap.f90:317-318): the sums are the sums of
    # the per-resample parameters, whatever their return code
    finite = np.isfinite(got["params"]).all(1)
    assert finite.all()
    assert np.allclose(got["sum_p"], got["params"].sum(0), rtol=1e-9)


@pytest.mark.gpu
def test_c5_ensemble_matches_the_reference_stream(oracle):
    """North-star check 2 on the stress config: the production path (counter-based generator) against the
    oracle's serial RANLUX+polar stream.  Means are meaningless here (heavy tails: sd 1e16), so the
    comparison is distributional: the non-convergence fraction within 3 binomial standard errors, and a
    two-sample Kolmogorov-Smirnov test on every parameter."""
    import scipy.stats

    import fitter_mc_b200 as fmc

    model, x, y, err = datasets.dataset("c5")
    start = datasets.C5_FITTED.copy()
    na, nb = 20000, 4000
    ctx = fmc.Context(0)
    try:
        ctx.set_data(model, x, y, err)
        ctx.launch(start, na, keep=True)
        got = ctx.finish()
    finally:
        ctx.close()
    oracle.initialise_rng()
    ref = oracle.bootstrap(model, x, y, err, start, nb, flags=_oracle_flags(oracle), nthreads=8)
    fa = got["rc_count"][7:11].sum() / na
    fb = ref["rc_count"][7:11].sum() / nb
    se = np.sqrt(fa * (1 - fa) / na + fb * (1 - fb) / nb)
    print("\n[c5 stress ensemble] codes 7-10: gpu %.4f (n=%d)  reference stream %.4f (n=%d)  diff %.2f se"
          % (fa, na, fb, nb, (fa - fb) / se))
    assert fa >= NONCONV_MIN and abs(fa - fb) <= 3.5 * se
    for k in range(4):
        ks = scipy.stats.ks_2samp(got["params"][:, k], ref["params"][:, k])
        print("   parameter %d: KS statistic %.4f p-value %.3f" % (k, ks.statistic, ks.pvalue))
        assert ks.pvalue > 1e-3
