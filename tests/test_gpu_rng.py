"""Device-side tests of the counter-based Gaussian generator that replaces the reference's serial
RANLUX + polar stream (rang, utils.f90:345-364; RANLUX, ranlux.f:116-152) -- csrc/rng.cuh:
Philox4x32-10 keyed by (seed, resample id, data point / 4) and Box-Muller in FP32 on the SFU.

  * raw bits: the published Random123 known-answer vectors, computed ON THE GPU through a debug
    entry point of the C ABI (the host image is checked in tests/test_host_emul.py);
  * distribution: 1e8 deviates exactly as the fit kernels use them (fmc_generate_offsets) against the
    normal distribution -- Kolmogorov-Smirnov distance on a 2^16-bin CDF, moments, tail masses beyond
    4 and 5 sigma (where an FP32 Box-Muller could go wrong), the extreme value;
  * independence: lag-1 correlations between neighbouring data points of a resample, between the same
    data point of neighbouring resamples, and between the members of a Box-Muller pair (z, z^2).
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fmc():
    import fitter_mc_b200

    fitter_mc_b200.lib()
    return fitter_mc_b200


def test_philox_known_answers_on_the_device(fmc):
    ck = np.array([[0, 0, 0, 0, 0, 0],
                   [0xffffffff] * 6,
                   [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0]], dtype=np.uint32)
    out = fmc.debug_philox4x32(ck)
    assert out[0].tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert out[1].tolist() == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert out[2].tolist() == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    # the counter layout of rng.cuh: (point/4, 0, id_lo, id_hi), key (seed_lo, seed_hi) -- the four
    # deviates of block b of resample id come from exactly that block
    seed, rid, blk = 31415, (5 << 32) + 77, 3
    words = fmc.debug_philox4x32(np.array([[blk, 0, rid & 0xffffffff, rid >> 32, seed & 0xffffffff, seed >> 32]],
                                          dtype=np.uint32))[0]
    z = fmc.generate_offsets(seed, rid, 1, 16)[0, 4 * blk:4 * blk + 4]
    u = (words[[0, 2]].astype(np.float64) + 0.5) * 2.0 ** -32
    ang = 2 * np.pi * words[[1, 3]].astype(np.float64) * 2.0 ** -32
    r = np.sqrt(-2 * np.log(u))
    want = np.array([r[0] * np.cos(ang[0]), r[0] * np.sin(ang[0]), r[1] * np.cos(ang[1]), r[1] * np.sin(ang[1])])
    assert np.allclose(z, want, rtol=2e-5, atol=2e-6), (z, want)     # FP32 / SFU evaluation of Box-Muller


def _phi(x):
    return 0.5 * (1.0 + np.vectorize(math.erf)(x / math.sqrt(2.0)))


def test_normals_distribution_tails_and_correlations_on_1e8_deviates(fmc):
    ndata, chunk, nchunk = 1000, 10000, 10          # 1e8 deviates in 10 chunks of 1e7
    nbins, lo, hi = 1 << 16, -8.0, 8.0
    hist = np.zeros(nbins, dtype=np.int64)
    n = 0
    s1 = s2 = s3 = s4 = 0.0
    tail4 = tail5 = 0
    zmax = 0.0
    c_point = c_resample = c_pair = c_sq = 0.0
    n_point = n_resample = n_pair = 0
    prev_last = None
    for c in range(nchunk):
        z = fmc.generate_offsets(31415, c * chunk, chunk, ndata)
        assert np.isfinite(z).all()
        hist += np.histogram(z, bins=nbins, range=(lo, hi))[0]
        n += z.size
        s1 += z.sum()
        z2 = z * z
        s2 += z2.sum()
        s3 += (z2 * z).sum()
        s4 += (z2 * z2).sum()
        a = np.abs(z)
        tail4 += int((a > 4).sum())
        tail5 += int((a > 5).sum())
        zmax = max(zmax, float(a.max()))
        c_point += float((z[:, 1:] * z[:, :-1]).sum())       # neighbouring data points of one resample
        n_point += z[:, 1:].size
        c_resample += float((z[1:, :] * z[:-1, :]).sum())    # same data point, neighbouring resamples
        n_resample += z[1:, :].size
        if prev_last is not None:
            c_resample += float((z[0] * prev_last).sum())
            n_resample += ndata
        prev_last = z[-1].copy()
        c_pair += float((z[:, 0::2] * z[:, 1::2]).sum())      # the two members of a Box-Muller pair
        c_sq += float(((z2[:, 0::2] - 1) * (z2[:, 1::2] - 1)).sum())
        n_pair += z[:, 0::2].size
    assert n == 100_000_000
    mean, var = s1 / n, s2 / n - (s1 / n) ** 2
    skew, kurt = s3 / n, s4 / n
    # Kolmogorov-Smirnov distance at the bin edges
    edges = np.linspace(lo, hi, nbins + 1)
    cdf = np.cumsum(hist) / n
    D = float(np.max(np.abs(cdf - _phi(edges[1:]))))
    p4, p5 = 2 * (1 - 0.5 * (1 + math.erf(4 / math.sqrt(2)))), 2 * (1 - 0.5 * (1 + math.erf(5 / math.sqrt(2))))
    r_point, r_res, r_pair, r_sq = c_point / n_point, c_resample / n_resample, c_pair / n_pair, c_sq / n_pair / 2.0
    print("\n[rng] n=%d mean %.2e var-1 %.2e skew %.2e kurt-3 %.2e ; KS D %.2e (1%% critical %.2e) ; |z|>4: %d "
          "(expect %.0f) |z|>5: %d (expect %.1f) max|z| %.3f ; lag-1 corr: points %.2e resamples %.2e pair %.2e "
          "pair(z^2) %.2e (1 sigma = %.1e)"
          % (n, mean, var - 1, skew, kurt - 3, D, 1.63 / math.sqrt(n), tail4, n * p4, tail5, n * p5, zmax, r_point,
             r_res, r_pair, r_sq, 1 / math.sqrt(n_point)))
    assert abs(mean) <= 4 / math.sqrt(n)
    assert abs(var - 1) <= 4 * math.sqrt(2.0 / n)
    assert abs(skew) <= 4 * math.sqrt(15.0 / n) and abs(kurt - 3) <= 4 * math.sqrt(96.0 / n)
    assert D <= 1.63 / math.sqrt(n)                              # 1 % critical value of the KS test
    assert abs(tail4 - n * p4) <= 4 * math.sqrt(n * p4)
    assert abs(tail5 - n * p5) <= 4 * math.sqrt(n * p5)
    assert 5.3 < zmax <= 6.77                                    # sqrt(-2 ln 2^-33) = 6.764 is the generator's bound
    for r, m in ((r_point, n_point), (r_res, n_resample), (r_pair, n_pair), (r_sq, n_pair)):
        assert abs(r) <= 4 / math.sqrt(m)
