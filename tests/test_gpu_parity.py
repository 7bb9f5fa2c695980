"""Parity of the CUDA path (through the C ABI of libfmc_b200.so) against the CPU oracle.

Bar (DESIGN.md section "Parity"): the work is FP64 with a finite-difference Jacobian, so the
gate is a tolerance, not bit equality.  BASELINE.json's north star asks for per-fit agreement
to rtol 1e-9 on identical injected resamples.  tests/test_parity_floor.py shows that the
reference algorithm itself, compiled at the reference's own default flags (-Ofast
-march=native) versus strict IEEE, differs per fit by a median of 1e-9..4e-9, a 99th
percentile of 5e-8 and a maximum of 7e-8 (forward-difference noise decides whether the last
noise-level step of the second nl2sno call is accepted), so 1e-9 is below the reference's own
reproducibility floor.  profiles/r2_parity_attribution.txt attributes the tail fit by fit: fits whose
solver takes the same decisions as the oracle's agree to 2e-9 (1e-16 with exact derivatives on both
sides); every fit beyond 1e-8 has a flipped decision in the noise-level last step of the second
call (stop test, or keep / restore x in assess, toms573.f90:1905-1948), with or without
perturbation arithmetic.  The tolerances asserted here are the measured floor:
   median rel. deviation <= 5e-9, max rel. deviation <= 1e-7, max deviation <= 1e-4 sigma,
and the measured distribution (with the fraction within 1e-9) is printed.
"""
import numpy as np
import pytest

import datasets

pytestmark = pytest.mark.gpu

REL_MEDIAN, REL_MAX, SIGMA_MAX = 5e-9, 1e-7, 1e-4


@pytest.fixture(scope="module")
def fmc():
    import fitter_mc_b200

    fitter_mc_b200.lib()
    return fitter_mc_b200


def _start(oracle, model, x, y, err):
    info = oracle.model_info(model)
    start, _ = oracle.fit_data(model, x, y, err, info["params_init"])
    return start


def _compare(name, got, want):
    rel = (np.abs(got - want) / np.abs(want)).max(1)
    sig = (np.abs(got - want) / want.std(0)).max()
    print("\n[parity %s] n=%d rel dev: median %.2e p99 %.2e max %.2e ; max %.2e sigma ; frac<=1e-9: %.3f"
          % (name, len(got), np.median(rel), np.quantile(rel, 0.99), rel.max(), sig, (rel <= 1e-9).mean()))
    assert np.median(rel) <= REL_MEDIAN
    assert rel.max() <= REL_MAX
    assert sig <= SIGMA_MAX


@pytest.mark.parametrize("name", ["c1", "c2p3", "c2", "c4", "c5mild"])
def test_unresampled_fit_matches_oracle(fmc, oracle, name):
    """bootstrap.f90:288-291 on the GPU (fmc_fit_once) vs the oracle's fit_data."""
    model, x, y, err = datasets.dataset(name)
    info = fmc.initialise_model(model)
    oinfo = oracle.model_info(model)
    assert info["nparam"] == oinfo["nparam"] and info["param_name"] == oinfo["param_name"]
    assert np.array_equal(info["params_init"], oinfo["params_init"])
    got, rc = fmc.fit_data(x, y, err, info["params_init"], model)
    want, cnt = oracle.fit_data(model, x, y, err, oinfo["params_init"])
    assert np.allclose(got, want, rtol=2e-7), (got, want)
    assert rc[0] in (3, 4, 5, 6) and rc[1] in (3, 4, 5, 6)


@pytest.mark.parametrize("name,nres", [("c1", 4096), ("c2p3", 1024), ("c2", 1024), ("c5mild", 1024), ("c4", 1024)])
def test_injected_resamples_per_fit_parity(fmc, oracle, name, nres):
    """Identical injected offsets to both codes: per-fit parameters, properties, moment sums."""
    model, x, y, err = datasets.dataset(name)
    start = _start(oracle, model, x, y, err)
    z = np.random.default_rng(2024).normal(size=(nres, len(x)))
    got = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=nres, params_init=start, initial_fit=False,
                                   z_inject=z, per_resample=True)
    want = oracle.bootstrap(model, x, y, err, start, nres, z=z, nthreads=8)
    _compare(name, got["params"], want["params"])
    assert np.allclose(got["props"], want["props"], rtol=1e-6)
    # moment sums (bootstrap.f90:317-318) and the final mean / Bessel-corrected error (328-341)
    assert np.allclose(got["sum_p"], want["sum_p"], rtol=1e-9)
    assert np.allclose(got["sum_p2"], want["sum_p2"], rtol=1e-9)
    assert np.allclose(got["sum_prop"], want["sum_prop"], rtol=1e-9)
    m, s = oracle.finalise_moments(nres, want["sum_p"], want["sum_p2"])
    assert np.allclose(got["params_opt"], m, rtol=1e-9) and np.allclose(got["err_params_opt"], s, rtol=1e-5)
    # the device sums must equal the sums of the per-resample outputs it reported
    assert np.allclose(got["sum_p"], got["params"].sum(0), rtol=1e-12)
    assert np.allclose(got["sum_p2"], (got["params"] ** 2).sum(0), rtol=1e-12)
    # return codes: same histogram up to the noise-level last step of the second nl2sno call
    # (whether that step is accepted, rejected or called "false convergence" is decided by the
    # sign of a rounding-level change of f; see tests/test_parity_floor.py)
    assert got["rc_count"].sum() == nres
    assert np.abs(got["rc_count"] - want["rc_count"]).sum() <= max(8, 0.35 * nres)
    assert abs(int(got["rc_count"][3:7].sum()) - int(want["rc_count"][3:7].sum())) <= max(2, 0.01 * nres)


def test_counter_generator_path_matches_oracle_on_dumped_offsets(fmc, oracle):
    """The production path (offsets regenerated on chip from the counter-based generator) fits
    exactly the data sets that fmc_generate_offsets reports, and the oracle agrees on them."""
    model, x, y, err = datasets.dataset("c2")
    start = _start(oracle, model, x, y, err)
    nres, seed, first = 512, 31415, 1000
    z = fmc.generate_offsets(seed, first, nres, len(x))
    assert abs(z.mean()) < 4 / np.sqrt(z.size) and abs(z.std() - 1) < 4 / np.sqrt(2 * z.size)
    got = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=nres, seed=seed, first_counter=first,
                                   params_init=start, initial_fit=False, per_resample=True)
    inj = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=nres, params_init=start, initial_fit=False,
                                   z_inject=z, per_resample=True)
    assert np.array_equal(got["params"], inj["params"])  # same kernel arithmetic, same offsets
    want = oracle.bootstrap(model, x, y, err, start, nres, z=z, nthreads=8)
    _compare("c2/generator", got["params"], want["params"])


def test_resample_ids_not_launch_geometry_decide_the_fits(fmc, oracle):
    """Counter-based ids: splitting [0,N) over launches (or GPUs/ranks) cannot change a fit."""
    model, x, y, err = datasets.dataset("c1")
    start = _start(oracle, model, x, y, err)
    n = 3000
    whole = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=n, params_init=start, initial_fit=False,
                                     per_resample=True)
    a = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=1234, first_counter=0, params_init=start,
                                 initial_fit=False, per_resample=True)
    b = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=n - 1234, first_counter=1234, params_init=start,
                                 initial_fit=False, per_resample=True)
    assert np.array_equal(whole["params"], np.vstack([a["params"], b["params"]]))
    assert np.allclose(whole["sum_p"], a["sum_p"] + b["sum_p"], rtol=1e-12)
    assert np.allclose(whole["sum_p2"], a["sum_p2"] + b["sum_p2"], rtol=1e-12)
    again = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=n, params_init=start, initial_fit=False,
                                     per_resample=True)
    assert np.array_equal(whole["params"], again["params"])


@pytest.mark.parametrize("name,na,nb", [("c1", 200000, 20000), ("c2", 200000, 6000), ("c2p3", 100000, 6000),
                                        ("c5mild", 100000, 6000), ("c4", 40000, 800)])
def test_ensemble_matches_reference_stream_statistics(fmc, oracle, name, na, nb):
    """North-star check 2 on every config: ensemble means and errors of the production path
    (counter-based generator, na resamples) vs the reference's own serial RANLUX+polar stream
    (different random numbers, nb resamples; bootstrap.f90:309-341), within 3 combined standard errors.
    (The stress config C5 is compared distributionally: tests/test_c5_stress.py.)"""
    model, x, y, err = datasets.dataset(name)
    start = _start(oracle, model, x, y, err)
    got = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=na, params_init=start, initial_fit=False)
    oracle.initialise_rng()
    ref = oracle.bootstrap(model, x, y, err, start, nb, per_resample=False, nthreads=8)
    rm, rs = oracle.finalise_moments(nb, ref["sum_p"], ref["sum_p2"])
    gm, gs = got["params_opt"], got["err_params_opt"]
    zm = (gm - rm) / np.sqrt(gs ** 2 / na + rs ** 2 / nb)
    zs = (gs - rs) / (rs * np.sqrt(0.5 / na + 0.5 / nb))
    print("\n[ensemble %s] gpu n=%d vs RANLUX-stream oracle n=%d: mean deviations %s se ; error-bar deviations %s se"
          % (name, na, nb, np.round(zm, 2).tolist(), np.round(zs, 2).tolist()))
    assert (np.abs(zm) <= 3.0).all(), (gm, rm)
    assert (np.abs(zs) <= 3.5).all(), (gs, rs)
    qm, qs = oracle.finalise_moments(nb, ref["sum_prop"], ref["sum_prop2"])
    assert (np.abs(got["props_opt"] - qm) <= 3 * np.sqrt(got["err_props_opt"] ** 2 / na + qs ** 2 / nb)).all()


def test_edge_cases_and_error_codes(fmc, oracle):
    model, x, y, err = datasets.dataset("c1")
    start = _start(oracle, model, x, y, err)
    # ragged sizes: ndata not a multiple of the 4-point generator block, down to ndata = nparam
    for n in (39, 38, 37, 5, 3):
        z = np.random.default_rng(n).normal(size=(16, n)) * (0.0 if n == 3 else 1.0)
        got = fmc.monte_carlo_fit_data(x[:n], y[:n], err[:n], model, nrand_points=16, params_init=start,
                                       initial_fit=False, z_inject=z, per_resample=True)
        want = oracle.bootstrap(model, x[:n], y[:n], err[:n], start, 16, z=z)
        if n > 10:
            assert np.allclose(got["params"], want["params"], rtol=1e-6), n
        else:
            # 3 parameters from 3-5 points: the valley is nearly flat and where the iteration stops in
            # it is decided at rounding level, so compare the objective reached, not the parameters
            for i in range(16):
                fg = (oracle.madr(model, x[:n], y[:n] + z[i] * err[:n], err[:n], got["params"][i]) ** 2).sum()
                fw = (oracle.madr(model, x[:n], y[:n] + z[i] * err[:n], err[:n], want["params"][i]) ** 2).sum()
                assert abs(fg - fw) <= 1e-3 * max(1.0, fw), (n, i, fg, fw)
    # empty run: zero resamples is a no-op, not an error
    got = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=0, params_init=start, initial_fit=False)
    assert got["rc_count"].sum() == 0 and (got["sum_p"] == 0).all()
    # a single resample (the reference's nrand_points=1 branch gives zero error bars)
    got = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=1, params_init=start, initial_fit=False)
    assert (got["err_params_opt"] == 0).all()
    # argument errors follow the C-ABI convention (negative code -> exception here)
    with pytest.raises(fmc.FmcError):
        fmc.monte_carlo_fit_data(x[:2], y[:2], err[:2], model, nrand_points=4, params_init=start, initial_fit=False)
    bad = err.copy()
    bad[3] = 0.0  # "Found a non-positive error bar", bootstrap.f90:134-138
    with pytest.raises(fmc.FmcError):
        fmc.monte_carlo_fit_data(x, y, bad, model, nrand_points=4, params_init=start, initial_fit=False)
    with pytest.raises(fmc.FmcError):
        fmc.initialise_model(99)


def test_non_converging_fits_are_averaged_like_any_other(fmc, oracle):
    """fit_data never inspects nl2sno's return code (bootstrap.f90:215-226): a start far from the
    minimum with huge noise gives limit / false-convergence codes that still enter the sums."""
    model = 3
    x, y, err = datasets.c5(n=200, rel=0.6, floor=0.5)
    nres = 512
    z = np.random.default_rng(5).normal(size=(nres, len(x)))
    start = np.array([1.0, 1.0, 1.0, 1.0])
    got = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=nres, params_init=start, initial_fit=False,
                                   z_inject=z, per_resample=True)
    want = oracle.bootstrap(model, x, y, err, start, nres, z=z, nthreads=8)
    assert got["rc_count"].sum() == nres
    assert np.isfinite(got["sum_p"]).all() == np.isfinite(want["sum_p"]).all()
    same = np.isclose(got["params"], want["params"], rtol=1e-5).all(1)
    print("\n[stress] rc histogram gpu %s oracle %s ; fits agreeing to 1e-5: %.3f"
          % (got["rc_count"][3:11], want["rc_count"][3:11], same.mean()))
    assert same.mean() > 0.9


def test_in_process_multi_gpu_split_equals_one_gpu(fmc, oracle):
    """fmc_bootstrap_run(ngpus=G): contiguous id ranges on G devices, one host thread each,
    accumulator blocks added on the host.  Same fits, sums equal up to summation order."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    model, x, y, err = datasets.dataset("c1")
    start = _start(oracle, model, x, y, err)
    n = 20001
    one = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=n, ngpus=1, params_init=start, initial_fit=False,
                                   per_resample=True)
    two = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=n, ngpus=2, params_init=start, initial_fit=False,
                                   per_resample=True)
    assert np.array_equal(one["params"], two["params"])
    assert np.allclose(one["sum_p"], two["sum_p"], rtol=1e-12) and np.allclose(one["sum_p2"], two["sum_p2"], rtol=1e-12)
    assert np.array_equal(one["rc_count"], two["rc_count"])


@pytest.mark.parametrize("name,nres", [("c2", 2048), ("c1", 4096), ("c4", 512)])
def test_parity_tail_on_the_gpu_is_a_decision_flip(fmc, oracle, name, nres):
    """Attribution of the per-fit deviations ON THE DEVICE (the fit-by-fit decision traces of
    tools/parity_attribution.py need the host image of the solver): the kernels report, per fit and nl2sno call,
    return code, function evaluations, gradient evaluations and iterations -- the oracle does too.  A fit whose
    eight counters equal the oracle's took the same decisions; the others flipped one somewhere.  The
    deviations beyond the forward-difference noise all belong to the second group, and the flips sit in the
    second call (the noise-level last steps, toms573.f90:1905-1948)."""
    model, x, y, err = datasets.dataset(name)
    start = _start(oracle, model, x, y, err)
    z = np.random.default_rng(4242).normal(size=(nres, len(x)))
    import torch

    ctx = fmc.Context(0)
    try:
        ctx.set_data(model, x, y, err)
        zdev = torch.from_numpy(z).cuda()
        torch.cuda.synchronize()
        ctx.launch(start, nres, z_inject_dev=zdev.data_ptr(), keep=True)
        got = ctx.finish()
    finally:
        ctx.close()
    want = oracle.bootstrap(model, x, y, err, start, nres, z=z, nthreads=8)
    rel = (np.abs(got["params"] - want["params"]) / np.abs(want["params"])).max(1)
    gc, oc = got["counters"][:, :, :4], want["counters"][:, :, :4]
    same = (gc == oc).reshape(nres, -1).all(1)
    first_call_differs = (gc[:, 0] != oc[:, 0]).any(1)
    rc_only = ~same & (gc[:, :, 1:] == oc[:, :, 1:]).reshape(nres, -1).all(1)
    print("\n[attribution %s] n=%d: same history as the oracle %.3f (median %.2e, max %.2e); flipped decision %.3f "
          "(median %.2e, max %.2e), of which already in the first call %.3f, return code only %.3f; "
          "fits beyond 5e-8: %d, all flipped: %s"
          % (name, nres, same.mean(), np.median(rel[same]), rel[same].max(), (~same).mean(),
             np.median(rel[~same]) if (~same).any() else 0.0, rel[~same].max() if (~same).any() else 0.0,
             first_call_differs[~same].mean() if (~same).any() else 0.0, rc_only[~same].mean() if (~same).any() else 0.0,
             (rel > 5e-8).sum(), bool(((rel > 5e-8) <= ~same).all())))
    assert same.mean() >= 0.2
    assert rel[same].max() <= 5e-8                  # same decisions: forward-difference noise only
    assert ((rel > 5e-8) <= ~same).all()            # the tail is made of flipped decisions
    assert first_call_differs[~same].mean() <= 0.5 if (~same).any() else True
