"""The N>1 path on CPU: two gloo ranks each fit a disjoint contiguous range of resample ids
(with the oracle standing in for the GPU kernel), all-reduce their accumulator blocks, and
decode them with the library's own fmc_decode_accumulator.  The result must equal the
single-rank result up to summation order (<= 1e-12 relative)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import datasets

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, nres, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import oracle_lib as oracle
    import fitter_mc_b200 as fmc
    from fitter_mc_b200 import distributed as fd

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model, x, y, err = datasets.dataset("c1")
    start, _ = oracle.fit_data(model, x, y, err, oracle.model_info(model)["params_init"])
    z = np.random.default_rng(77).normal(size=(nres, len(x)))   # "global" resamples, same on every rank
    lo, hi = fd.shard_range(nres, rank, world)
    out = oracle.bootstrap(model, x, y, err, start, hi - lo, z=z[lo:hi])
    acc = fd.encode_accumulator(out["params"], out["props"], start, fmc.derived_properties(model, start),
                                rc=out["counters"][:, 1, 0], counters=out["counters"])
    t = torch.from_numpy(acc)
    fd.allreduce_accumulator(t)
    dec = fmc.decode_accumulator(model, start, t.numpy())
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi, **{k: np.asarray(v) for k, v in dec.items()})
    dist.destroy_process_group()


def test_shard_range_partitions_the_ids():
    from fitter_mc_b200 import distributed as fd

    for n in (0, 1, 7, 1000, 10 ** 8):
        for world in (1, 2, 3, 8):
            r = [fd.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_two_gloo_ranks_equal_one_rank(tmp_path, oracle):
    import fitter_mc_b200 as fmc

    nres, world = 600, 2
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, nres, str(tmp_path)), nprocs=world, join=True)
    model, x, y, err = datasets.dataset("c1")
    start, _ = oracle.fit_data(model, x, y, err, oracle.model_info(model)["params_init"])
    z = np.random.default_rng(77).normal(size=(nres, len(x)))
    one = oracle.bootstrap(model, x, y, err, start, nres, z=z)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert (int(r0["lo"]), int(r0["hi"]), int(r1["lo"]), int(r1["hi"])) == (0, 300, 300, 600)
    for k in ("sum_p", "sum_p2", "sum_prop", "sum_prop2", "rc_count", "totals", "nfit"):
        assert np.array_equal(r0[k], r1[k])           # every rank decodes the same reduced block
    assert int(r0["nfit"]) == nres
    assert np.allclose(r0["sum_p"], one["sum_p"], rtol=1e-12)
    assert np.allclose(r0["sum_p2"], one["sum_p2"], rtol=1e-12)
    assert np.allclose(r0["sum_prop"], one["sum_prop"], rtol=1e-12)
    assert np.allclose(r0["sum_prop2"], one["sum_prop2"], rtol=1e-12)
    assert np.array_equal(r0["rc_count"], one["rc_count"])
    m, s = fmc.finalise_moments(nres, r0["sum_p"], r0["sum_p2"])
    om, os_ = oracle.finalise_moments(nres, one["sum_p"], one["sum_p2"])
    assert np.allclose(m, om, rtol=1e-13) and np.allclose(s, os_, rtol=1e-9)
