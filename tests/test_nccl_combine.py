"""The single collective of the path, inside the C ABI: the moment
accumulators of the GPUs of one box are combined by one ncclAllReduce over NVLink, replacing the
OpenMP REDUCTION(+) of bootstrap.f90:310-311 across GPUs.  NCCL is bound at run time (dlopen), so
the CPU part only checks that it loads; the GPU parts need two devices
(`gpurun --gpus 2 -- python -m pytest tests/test_nccl_combine.py -m gpu`)."""
import threading

import numpy as np
import pytest

import datasets


def test_nccl_loads_and_argument_checks():
    import fitter_mc_b200 as fmc

    L = fmc.lib()
    assert fmc.nccl_available() >= 1                        # libnccl.so.2 is part of the image
    assert fmc.last_combine_method() in ("none", "nccl", "host")
    assert L.fmc_nccl_unique_id(None) == -4                  # FMC_ERR_BAD_ARG
    assert L.fmc_nccl_init_rank(None, b"\0" * 128, 2, 0) == -4
    assert L.fmc_allreduce(None) == -4
    assert "NCCL" in L.fmc_strerror(-8).decode()


def _two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


@pytest.mark.gpu
def test_one_call_multi_gpu_run_combines_with_nccl(oracle):
    """fmc_bootstrap_run(ngpus=2): the path the Fortran drop-in uses.  Same fits as on one GPU, sums
    equal up to the order of summation, and the combine really was the NCCL all-reduce."""
    import fitter_mc_b200 as fmc

    _two_gpus()
    model, x, y, err = datasets.dataset("c2")
    start, _ = oracle.fit_data(model, x, y, err, oracle.model_info(model)["params_init"])
    n = 300001
    one = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=n, ngpus=1, params_init=start, initial_fit=False)
    assert fmc.last_combine_method() == "none"
    two = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=n, ngpus=2, params_init=start, initial_fit=False)
    assert fmc.last_combine_method() == "nccl"
    for k in ("sum_p", "sum_p2", "sum_prop", "sum_prop2"):
        assert np.allclose(one[k], two[k], rtol=1e-12), k
    assert np.array_equal(one["rc_count"], two["rc_count"]) and np.array_equal(one["totals"], two["totals"])
    # binned distributions: integer counts, combined by the same collective, exactly additive
    b1 = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=50000, ngpus=1, params_init=start, initial_fit=False,
                                  histogram_bins=64)
    b2 = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=50000, ngpus=2, params_init=start, initial_fit=False,
                                  histogram_bins=64, histogram_range=(b1["hist_lo"], b1["hist_hi"]))
    assert fmc.last_combine_method() == "nccl"
    assert np.array_equal(b1["hist_counts"], b2["hist_counts"])


@pytest.mark.gpu
def test_rank_per_gpu_allreduce_through_the_c_abi(oracle):
    """One rank per GPU with the library's own communicator (fmc_nccl_unique_id / fmc_nccl_init_rank /
    fmc_allreduce): two host threads stand in for two processes."""
    import fitter_mc_b200 as fmc

    _two_gpus()
    model, x, y, err = datasets.dataset("c1")
    start, _ = oracle.fit_data(model, x, y, err, oracle.model_info(model)["params_init"])
    n = 40000
    uid = fmc.nccl_unique_id()
    out, errors = [None, None], []

    def rank(r):
        try:
            ctx = fmc.Context(r)
            ctx.set_data(model, x, y, err)
            ctx.nccl_init_rank(uid, 2, r)
            ctx.launch(start, n // 2, first_counter=r * (n // 2))
            ctx.allreduce()
            ctx.wait()
            out[r] = ctx
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    th = [threading.Thread(target=rank, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    assert not errors, errors
    whole = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=n, ngpus=1, params_init=start, initial_fit=False)
    for r in range(2):
        got = out[r].read_accumulator()
        dec = fmc.decode_accumulator(model, start, got)
        assert dec["nfit"] == n                              # every rank holds the total
        assert np.allclose(dec["sum_p"], whole["sum_p"], rtol=1e-12)
        assert np.allclose(dec["sum_p2"], whole["sum_p2"], rtol=1e-12)
        assert np.array_equal(dec["rc_count"], whole["rc_count"])
        out[r].close()
