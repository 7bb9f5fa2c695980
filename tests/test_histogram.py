"""Binned distributions of the fitted parameters and properties: the
device-side replacement for the reference's histogram_<prop>.dat dumps (bootstrap.f90:301-308,
319-325, 348-355).

CPU part: the host image of the binning rule, the quantile routine of the C ABI and the
integer all-reduce over two gloo ranks.  GPU part: counts from the CUDA path equal a re-binning
of the per-resample results exactly, and equal the binning of the ORACLE's results except for
values within the parity tolerance of a bin edge."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import datasets

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def fmc():
    import fitter_mc_b200

    fitter_mc_b200.lib()
    return fitter_mc_b200


def test_bin_values_is_numpy_histogram_with_outflow_columns():
    from fitter_mc_b200 import distributed as fd

    rng = np.random.default_rng(3)
    v = np.column_stack([rng.normal(1.0, 0.3, 20000), rng.exponential(2.0, 20000)])
    v[17, 0] = np.nan
    v[18, 1] = np.inf
    v[19, 1] = -np.inf
    lo, hi, nb = np.array([0.2, 0.5]), np.array([1.9, 7.0]), 64
    got = fd.bin_values(v, nb, lo, hi)
    assert got.shape == (2, nb + 2) and got.dtype == np.uint64
    assert np.all(got.sum(1) == len(v))                    # every fit lands somewhere
    for j in range(2):
        col = v[:, j]
        want, _ = np.histogram(col[np.isfinite(col)], bins=nb, range=(lo[j], hi[j]))
        # numpy closes the last bin on the right and may differ by rounding exactly at an edge
        assert np.abs(got[j, 1:-1].astype(np.int64) - want).sum() <= 2
        assert got[j, 0] == np.sum(col < lo[j]) + np.sum(np.isnan(col))
        assert got[j, -1] == np.sum(col >= hi[j])
    # the edges themselves: lo belongs to bin 1, hi to the overflow column
    e = fd.bin_values(np.array([[0.0], [1.0], [np.nextafter(1.0, 0.0)]]), 4, [0.0], [1.0])
    assert e[0].tolist() == [0, 1, 0, 0, 1, 1]


def test_quantiles_from_counts_match_the_sample_quantiles(fmc):
    from fitter_mc_b200 import distributed as fd

    rng = np.random.default_rng(11)
    v = rng.normal(3.0, 0.5, size=(400000, 1))
    lo, hi, nb = np.array([0.0]), np.array([6.0]), 1200
    counts = fd.bin_values(v, nb, lo, hi)[0]
    probs = np.array([0.001, 0.025, 0.16, 0.5, 0.84, 0.975, 0.999])
    q, outside = fmc.histogram_quantiles(counts, lo[0], hi[0], probs)
    assert outside == 0
    assert np.all(np.abs(q - np.quantile(v[:, 0], probs)) <= (hi[0] - lo[0]) / nb)   # within one bin width
    assert np.all(np.diff(q) > 0)
    # a range that is too narrow is reported, and the clipped quantiles sit on its ends
    c2 = fd.bin_values(v, 100, [2.9], [3.1])[0]
    q2, outside2 = fmc.histogram_quantiles(c2, 2.9, 3.1, [0.01, 0.5, 0.99])
    assert outside2 == int(c2[0] + c2[-1]) > 0.8 * len(v)
    assert q2[0] == 2.9 and q2[2] == 3.1 and abs(q2[1] - 3.0) < 0.01
    # empty histogram and bad probabilities give NaN, bad arguments an error
    q3, _ = fmc.histogram_quantiles(np.zeros(12, dtype=np.uint64), 0.0, 1.0, [0.5])
    assert np.isnan(q3[0])
    q4, _ = fmc.histogram_quantiles(counts, lo[0], hi[0], [-0.1, 1.5])
    assert np.isnan(q4).all()
    with pytest.raises(fmc.FmcError):
        fmc.histogram_quantiles(counts, 1.0, 1.0, [0.5])


def test_pilot_range_covers_the_sample(fmc):
    rng = np.random.default_rng(5)
    v = np.column_stack([rng.normal(2.0, 0.1, 5000), np.full(5000, 7.0)])
    v[3, 0] = np.nan                                        # a failed fit in the pilot is ignored
    lo, hi = fmc.histogram_range_from_pilot(v, nsigma=8.0)
    assert lo[0] < 1.3 and hi[0] > 2.7 and lo[1] < 7.0 < hi[1] and np.all(hi > lo)


def _hist_worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    import oracle_lib as oracle
    from fitter_mc_b200 import distributed as fd

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model, x, y, err = datasets.dataset("c1")
    start, _ = oracle.fit_data(model, x, y, err, oracle.model_info(model)["params_init"])
    nres = 500
    z = np.random.default_rng(78).normal(size=(nres, len(x)))
    lo, hi = fd.shard_range(nres, rank, world)
    out = oracle.bootstrap(model, x, y, err, start, hi - lo, z=z[lo:hi])
    vals = np.column_stack([out["params"], out["props"]])
    rlo, rhi = np.load(os.path.join(out_dir, "range.npy"))
    counts = fd.bin_values(vals, 50, rlo, rhi)
    t = torch.from_numpy(counts.astype(np.int64))
    fd.allreduce_histogram(t)
    np.save(os.path.join(out_dir, "hist%d.npy" % rank), t.numpy())
    dist.destroy_process_group()


def test_two_gloo_ranks_sum_their_histograms_exactly(tmp_path, oracle):
    from fitter_mc_b200 import distributed as fd

    model, x, y, err = datasets.dataset("c1")
    start, _ = oracle.fit_data(model, x, y, err, oracle.model_info(model)["params_init"])
    nres = 500
    z = np.random.default_rng(78).normal(size=(nres, len(x)))
    one = oracle.bootstrap(model, x, y, err, start, nres, z=z)
    vals = np.column_stack([one["params"], one["props"]])
    rlo, rhi = vals.mean(0) - 3 * vals.std(0), vals.mean(0) + 3 * vals.std(0)   # some outflow on purpose
    np.save(tmp_path / "range.npy", np.stack([rlo, rhi]))
    port = 31500 + os.getpid() % 2000
    mp.spawn(_hist_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    h0, h1 = np.load(tmp_path / "hist0.npy"), np.load(tmp_path / "hist1.npy")
    want = fd.bin_values(vals, 50, rlo, rhi).astype(np.int64)
    assert np.array_equal(h0, h1) and np.array_equal(h0, want)
    assert np.all(h0.sum(1) == nres) and h0[:, 0].sum() + h0[:, -1].sum() > 0


def test_histogram_needs_a_gpu_context_and_valid_ranges(fmc):
    """No CPU fallback: without a device the context cannot be created at all."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test below")
    with pytest.raises(fmc.FmcError) as e:
        fmc.Context(0)
    assert e.value.code == -1


@pytest.mark.gpu
def test_device_histogram_equals_rebinning_of_per_resample_results(fmc, oracle):
    from fitter_mc_b200 import distributed as fd

    model, x, y, err = datasets.dataset("c1")
    info = fmc.initialise_model(model)
    start, _ = oracle.fit_data(model, x, y, err, info["params_init"])
    ctx = fmc.Context(0)
    ctx.set_data(model, x, y, err)
    # pilot: the range comes from per-resample results of a short run
    ctx.launch(start, 2000, keep=True)
    pilot = ctx.finish()
    lo, hi = fmc.histogram_range_from_pilot(np.column_stack([pilot["params"], pilot["props"]]), nsigma=4.0)
    nb, nres = 200, 30000
    with pytest.raises(fmc.FmcError):
        ctx.set_histogram(nb, hi, lo)                       # hi <= lo is refused
    ctx.set_histogram(nb, lo, hi)
    assert ctx.histogram_len() == (info["nparam"] + info["nprop"]) * (nb + 2)
    ctx.launch(start, nres, keep=True)
    out = ctx.finish()
    got = ctx.histogram()
    vals = np.column_stack([out["params"], out["props"]])
    assert np.array_equal(got, fd.bin_values(vals, nb, lo, hi))         # exact integer counts
    assert np.all(got.sum(1) == nres)
    assert got[:, 0].sum() + got[:, -1].sum() > 0                       # 4 sigma: some outflow
    # without the per-resample arrays (the large-N mode) the counts are the same
    ctx.launch(start, nres, keep=False)
    ctx.finish()
    assert np.array_equal(ctx.histogram(), got)
    # and they are additive over a split of the id range (what the N-GPU all-reduce relies on)
    ctx.launch(start, 12345, first_counter=0)
    ctx.finish()
    a = ctx.histogram()
    ctx.launch(start, nres - 12345, first_counter=12345)
    ctx.finish()
    assert np.array_equal(a + ctx.histogram(), got)
    # quantiles from the counts against the sample quantiles
    probs = np.array([0.025, 0.16, 0.5, 0.84, 0.975])
    for j in range(vals.shape[1]):
        q, outside = fmc.histogram_quantiles(got[j], lo[j], hi[j], probs)
        assert outside == int(got[j, 0] + got[j, -1])
        assert np.all(np.abs(q - np.quantile(vals[:, j], probs)) <= 1.5 * (hi[j] - lo[j]) / nb)
    # switching binning off again
    ctx.set_histogram(0)
    assert ctx.histogram_len() == 0
    ctx.launch(start, 100)
    ctx.finish()
    ctx.close()


@pytest.mark.gpu
def test_device_histogram_matches_binned_oracle_results(fmc, oracle):
    """Against the oracle on identical injected resamples: the counts may differ only where a
    value lies within the parity tolerance of a bin edge."""
    import torch
    from fitter_mc_b200 import distributed as fd

    model, x, y, err = datasets.dataset("c2")
    info = fmc.initialise_model(model)
    start, _ = oracle.fit_data(model, x, y, err, info["params_init"])
    nres, nb = 4000, 40
    z = np.random.default_rng(2024).normal(size=(nres, len(x)))
    want = oracle.bootstrap(model, x, y, err, start, nres, z=z, nthreads=8)
    wv = np.column_stack([want["params"], want["props"]])
    lo, hi = wv.mean(0) - 5 * wv.std(0), wv.mean(0) + 5 * wv.std(0)
    ctx = fmc.Context(0)
    ctx.set_data(model, x, y, err)
    ctx.set_histogram(nb, lo, hi)
    zd = torch.from_numpy(z).cuda()
    ctx.launch(start, nres, z_inject_dev=zd.data_ptr())
    ctx.finish()
    got = ctx.histogram().astype(np.int64)
    ref = fd.bin_values(wv, nb, lo, hi).astype(np.int64)
    assert np.all(got.sum(1) == nres)
    moved = np.abs(got - ref).sum(1) // 2           # fits that changed bin, per quantity
    print("\n[histogram vs oracle] fits in a different bin per quantity:", moved.tolist())
    assert moved.max() <= 2                         # expected ~ nres * nb * 1e-8 / 10 sigma: none
    ctx.close()
