"""The round schedule (csrc/round_kernels.cuh: solver steps and passes in separate kernels alternating
inside a CUDA graph WHILE node, fits compacted between rounds) against the refill schedule
(csrc/bootstrap_kernel.cuh) and the oracle.  Both schedules run the same pass_loop and the same
Nl2::resume per fit, so per-fit results must be IDENTICAL, bit for bit, whatever the pool size,
the chunking or the way the rounds are driven (graph or host loop)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import datasets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_schedule_argument_checks():
    import fitter_mc_b200 as fmc

    L = fmc.lib()
    assert L.fmc_set_schedule(None, 7) == -4
    assert L.fmc_set_schedule(None, 1) == 0 and L.fmc_get_schedule(None) == 1
    assert L.fmc_set_schedule(None, 0) == 0 and L.fmc_get_schedule(None) == 0


def _start(oracle, name, model, x, y, err):
    if name == "c5":
        return datasets.C5_FITTED.copy()
    return oracle.fit_data(model, x, y, err, oracle.model_info(model)["params_init"])[0]


def _run(fmc, schedule, model, x, y, err, start, nres, z=None, first=0, jac="fd"):
    import torch

    ctx = fmc.Context(0)
    try:
        ctx.set_data(model, x, y, err)
        ctx.set_schedule(schedule)
        ctx.set_jacobian_mode(jac)
        zdev = None
        if z is not None:
            zdev = torch.from_numpy(np.ascontiguousarray(z)).cuda()
            torch.cuda.synchronize()
        ctx.launch(start, nres, first_counter=first, z_inject_dev=zdev.data_ptr() if zdev is not None else None,
                   keep=True)
        out = ctx.finish()
        out["launches"] = ctx.launch_count()
        out["anomalies"] = ctx.anomalies()[2]
        return out
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,nres,inject", [("c1", 5000, False), ("c2", 3000, False), ("c2", 1024, True),
                                              ("c2p3", 2000, False), ("c5mild", 2000, False), ("c4", 600, False),
                                              ("c5", 3000, False)])
def test_rounds_and_refill_give_identical_fits(oracle, name, nres, inject):
    import fitter_mc_b200 as fmc

    model, x, y, err = datasets.dataset(name)
    start = _start(oracle, name, model, x, y, err)
    z = np.random.default_rng(7).normal(size=(nres, len(x))) if inject else None
    a = _run(fmc, "refill", model, x, y, err, start, nres, z)
    b = _run(fmc, "rounds", model, x, y, err, start, nres, z)
    c = _run(fmc, "rounds", model, x, y, err, start, nres, z)
    # the round schedule is deterministic: scheduling order never enters a fit
    assert np.array_equal(b["params"], c["params"]) and np.array_equal(b["counters"], c["counters"])
    # Refill and rounds run the same solver SOURCE, compiled differently (unrolled and inlined into the
    # refill kernel; rolled loops and out-of-line helpers in the solve kernel, which is short of
    # instruction cache, not registers): ptxas fuses a few multiply-adds differently, so a small share
    # of the fits differs in the last bits and -- where the last, noise-level step is decided by the sign of
    # a rounding error -- by the parity floor.  Everything else is identical.
    # (C5's fits run 30 ... 500 passes along flat valleys and are chaotic: a last-bit difference anywhere
    # changes the parameters a fit ends with -- between ANY two builds, tests/datasets.py -- so there only
    # the ensemble is compared here: return-code histogram and evaluation totals within 1 %; the per-fit
    # chi-squared gates against the oracle are in tests/test_c5_stress.py, for both schedules.)
    # (return codes, evaluation, gradient and iteration counts; the number of PASSES may differ: with many
    # parameters the round schedule evaluates the residual sum of a trial point first and its Jacobian only
    # when the step is accepted and the call goes on -- same decisions, one more pass per such step)
    same = (a["params"] == b["params"]).all(1) & (a["counters"][:, :, :4] == b["counters"][:, :, :4]).reshape(nres, -1).all(1)
    rel = (np.abs(a["params"] - b["params"]) / np.abs(a["params"])).max(1)
    print("\n[rounds %s] %d fits: %.4f bit-identical to the refill schedule, max rel dev of the others %.1e; "
          "kernels launched: refill %d, rounds %d" % (name, nres, same.mean(), rel.max(), a["launches"], b["launches"]))
    if name != "c5":
        assert same.mean() >= 0.99
        assert rel.max() <= 1e-7
        assert np.allclose(a["sum_p"], b["sum_p"], rtol=1e-9) and np.allclose(a["sum_p2"], b["sum_p2"], rtol=1e-9)
        assert np.allclose(a["sum_prop"], b["sum_prop"], rtol=1e-9)
    assert np.abs(a["rc_count"].astype(float) - b["rc_count"]).max() <= 0.01 * nres
    assert np.abs(a["totals"][:3].astype(float) / b["totals"][:3] - 1).max() <= 0.01
    assert np.allclose(b["sum_p"], b["params"].sum(0), rtol=1e-9)
    assert abs(a["anomalies"] - b["anomalies"]) <= 1


@pytest.mark.gpu
def test_rounds_analytic_mode_and_small_pool_and_host_driven_rounds(oracle):
    """Exact-Jacobian mode under the round schedule; a pool much smaller than the chunk (every slot is
    refilled many times); rounds driven by the host instead of the graph's WHILE node (FMC_ROUNDS_NO_GRAPH=1,
    the fallback for drivers without conditional graph nodes).  Each in a fresh process (environment)."""
    code = r"""
import sys, json, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import datasets, oracle_lib as oracle, fitter_mc_b200 as fmc
model, x, y, err = datasets.dataset("c2")
start = oracle.fit_data(model, x, y, err, oracle.model_info(model)["params_init"])[0]
out = {}
for sched in ("refill", "rounds"):
    for jac in ("fd", "analytic"):
        ctx = fmc.Context(0); ctx.set_data(model, x, y, err); ctx.set_schedule(sched); ctx.set_jacobian_mode(jac)
        ctx.launch(start, 40000, first_counter=123, keep=True); r = ctx.finish()
        out[sched + "/" + jac] = [r["params"].tobytes().hex()[:64], float(r["sum_p"][0]), int(r["counters"].sum()),
                                  __import__("hashlib").sha256(r["params"].tobytes()).hexdigest(), ctx.launch_count()]
        ctx.close()
print(json.dumps(out))
""" % dict(root=ROOT)
    res = {}
    for tag, env in (("graph", {}), ("small-pool", {"FMC_ROUND_WAVES": "1"}), ("host-driven", {"FMC_ROUNDS_NO_GRAPH": "1"})):
        e = dict(os.environ)
        e.update(env)
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        res[tag] = __import__("json").loads(p.stdout.strip().splitlines()[-1])
    ref = res["graph"]
    for tag, r in res.items():
        for jac in ("fd", "analytic"):
            # same fits, bit for bit, however the rounds are driven and however small the pool
            assert r["rounds/" + jac][3] == ref["rounds/" + jac][3], (tag, jac)
            assert r["rounds/" + jac][2] == ref["rounds/" + jac][2]
            # against the refill schedule: the same up to the few fits whose last bits differ (see above)
            assert abs(r["rounds/" + jac][1] / ref["refill/" + jac][1] - 1) <= 1e-9
            assert abs(r["rounds/" + jac][2] / ref["refill/" + jac][2] - 1) <= 1e-3
    assert ref["refill/fd"][3] != ref["refill/analytic"][3]
    print("\n[rounds] launches: %s" % {t: r["rounds/fd"][4] for t, r in res.items()})


@pytest.mark.gpu
@pytest.mark.parametrize("name,nres", [("c2", 4000), ("c4", 400), ("c5", 1500)])
def test_residual_first_rounds_give_the_same_fits_bit_for_bit(oracle, name, nres):
    """RoundArgs::spec: Jacobian with every trial point (1) or residual sum first (0, the default for more than
    six parameters).  Same solver build, same numbers read in the same order: identical fits, whatever the data."""
    import fitter_mc_b200 as fmc

    model, x, y, err = datasets.dataset(name)
    start = _start(oracle, name, model, x, y, err)
    out = {}
    old = os.environ.get("FMC_ROUND_SPECULATE")
    try:
        for spec in ("1", "0"):
            os.environ["FMC_ROUND_SPECULATE"] = spec
            out[spec] = _run(fmc, "rounds", model, x, y, err, start, nres, None)
    finally:
        if old is None:
            os.environ.pop("FMC_ROUND_SPECULATE", None)
        else:
            os.environ["FMC_ROUND_SPECULATE"] = old
    a, b = out["1"], out["0"]
    assert np.array_equal(a["params"], b["params"]) and np.array_equal(a["props"], b["props"])
    assert np.array_equal(a["counters"][:, :, :4], b["counters"][:, :, :4])
    assert np.array_equal(a["rc_count"], b["rc_count"]) and np.array_equal(a["totals"][:3], b["totals"][:3])
    pa, pb = a["counters"][:, :, 4].sum(1).mean(), b["counters"][:, :, 4].sum(1).mean()
    print("\n[rounds %s] residual first: %d fits identical; passes per fit %.2f -> %.2f (with a Jacobian: %.2f -> %.2f); "
          "kernels launched %d -> %d" % (name, nres, pa, pb, pa, pb - pa + 2.0, a["launches"], b["launches"]))
    assert pb >= pa


@pytest.mark.gpu
def test_rounds_per_fit_parity_with_the_oracle(oracle):
    import fitter_mc_b200 as fmc

    model, x, y, err = datasets.dataset("c2")
    start = _start(oracle, "c2", model, x, y, err)
    nres = 1024
    z = np.random.default_rng(2024).normal(size=(nres, len(x)))
    got = _run(fmc, "rounds", model, x, y, err, start, nres, z)
    want = oracle.bootstrap(model, x, y, err, start, nres, z=z, nthreads=8)
    rel = (np.abs(got["params"] - want["params"]) / np.abs(want["params"])).max(1)
    print("\n[rounds parity c2] median %.2e p99 %.2e max %.2e" % (np.median(rel), np.quantile(rel, 0.99), rel.max()))
    assert np.median(rel) <= 5e-9 and rel.max() <= 1e-7
    assert np.allclose(got["sum_p"], want["sum_p"], rtol=1e-9)


def test_schedule_setting_argument_checks():
    """No GPU needed: the setting is validated and kept per process for the one-call entry points."""
    import fitter_mc_b200 as fmc

    L = fmc.lib()
    try:
        for code in (0, 1, 2):          # FMC_SCHEDULE_REFILL, _ROUNDS, _AUTO
            assert L.fmc_set_schedule(None, code) == 0
            assert L.fmc_get_schedule(None) == code
        assert L.fmc_set_schedule(None, 3) == -4 and L.fmc_set_schedule(None, -1) == -4   # FMC_ERR_BAD_ARG
        assert L.fmc_get_schedule(None) == 2
    finally:
        L.fmc_set_schedule(None, 2)
