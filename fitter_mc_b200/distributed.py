"""Multi-GPU / multi-rank host logic: the path shards over independent resamples.

SURVEY.md section 8e: resamples are independent once the generator is counter-based, so rank
r of G takes a contiguous range of global resample ids and no data-path collective is needed;
the only exchange is ONE all-reduce (sum) of the accumulator block -- 2(p+nprop) moment sums
plus 21 counters, all doubles, < 1 KB -- which every rank then decodes identically.
torch.distributed is plumbing here (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_range(nresample, rank, world):
    """Contiguous split of [0, nresample): rank r gets [lo, hi)."""
    lo = nresample * rank // world
    hi = nresample * (rank + 1) // world
    return lo, hi


def accumulator_layout(nparam, nprop):
    """Offsets inside the accumulator block (bootstrap_kernel.cuh)."""
    nm = 2 * (nparam + nprop)
    return dict(sum_p=(0, nparam), sum_p2=(nparam, 2 * nparam), sum_prop=(2 * nparam, 2 * nparam + nprop),
                sum_prop2=(2 * nparam + nprop, nm), nfit=nm, rc=(nm + 1, nm + 17), nfcall=nm + 17,
                ngcall=nm + 18, niter=nm + 19, npass=nm + 20, length=nm + 21)


def encode_accumulator(params, props, params_start, props_start, rc=None, counters=None):
    """Host-side image of what the kernel accumulates for a set of finished fits: sums of the
    values SHIFTED by their value at params_start.  Used by the CPU tests of the N>1 path."""
    params = np.asarray(params, dtype=np.float64)
    props = np.asarray(props, dtype=np.float64).reshape(len(params), -1)
    p, q = params.shape[1], props.shape[1]
    lay = accumulator_layout(p, q)
    acc = np.zeros(lay["length"])
    dp = params - np.asarray(params_start)[None, :]
    dq = props - np.asarray(props_start)[None, :]
    acc[0:p] = dp.sum(0)
    acc[p:2 * p] = (dp * dp).sum(0)
    acc[2 * p:2 * p + q] = dq.sum(0)
    acc[2 * p + q:2 * p + 2 * q] = (dq * dq).sum(0)
    acc[lay["nfit"]] = len(params)
    if rc is not None:
        acc[lay["rc"][0]:lay["rc"][1]] = np.bincount(np.asarray(rc) & 15, minlength=16)
    if counters is not None:
        c = np.asarray(counters)
        acc[lay["nfcall"]], acc[lay["ngcall"]], acc[lay["niter"]] = c[..., 1].sum(), c[..., 2].sum(), c[..., 3].sum()
    return acc


def allreduce_accumulator(acc):
    """The single collective of the path.  `acc` is a torch tensor (CUDA with NCCL, CPU with gloo)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return acc


def bin_values(values, nbins, lo, hi):
    """Host image of the device binning (fmc::hist_bin, bootstrap_kernel.cuh): counts
    [ncol][nbins + 2] of `values` [nfit][ncol]; column 0 = underflow and NaN, nbins + 1 = overflow.
    The same three FP64 operations as the kernel, so the counts agree exactly."""
    v = np.asarray(values, dtype=np.float64)
    lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    scale = float(nbins) / (hi - lo)
    out = np.zeros((v.shape[1], nbins + 2), dtype=np.uint64)
    for j in range(v.shape[1]):
        t = (v[:, j] - lo[j]) * scale[j]
        b = np.zeros(len(t), dtype=np.int64)
        inside = (t >= 0.0) & (t < float(nbins))
        b[inside] = 1 + t[inside].astype(np.int64)
        b[t >= float(nbins)] = nbins + 1
        out[j] = np.bincount(b, minlength=nbins + 2).astype(np.uint64)
    return out


def allreduce_histogram(counts):
    """Sum the binned distributions over ranks.  `counts` is a torch int64 tensor (the device
    buffer handed to fmc_set_histogram_dev, or a CPU tensor with gloo).  Integer sums: the
    result is exact and independent of the rank order."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts
