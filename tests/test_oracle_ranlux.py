"""Pin the oracle's RANLUX/rang restatement on the reference's own known-answer vectors
(ranlux.f:399-473, the expected output of LUXTST -- the only golden vectors the reference ships)."""
import numpy as np


def _chk(got, want):
    assert np.allclose(got, want, rtol=0, atol=5.1e-9), (got, want)


def test_default_initialisation(oracle):
    oracle.rlux_reset_default()  # NOTYET=.TRUE.: seed 314159265, luxury level 3 (ranlux.f:399-407)
    _chk(oracle.ranlux(100)[:5], [0.53981817, 0.76155043, 0.06029940, 0.79600263, 0.30631220])
    _chk(oracle.ranlux(100)[:5], [0.43156743, 0.03774416, 0.24897110, 0.00147784, 0.90274453])


def test_luxury_level_0(oracle):
    oracle.rluxgo(0, 0)  # ranlux.f:409-415
    _chk(oracle.ranlux(100)[:5], [0.53981817, 0.76155043, 0.06029940, 0.79600263, 0.30631220])
    _chk(oracle.ranlux(100)[:5], [0.41538775, 0.05330932, 0.58195311, 0.91397446, 0.67034441])


def test_luxury_p389_seed1(oracle):
    oracle.rluxgo(389, 1)  # ranlux.f:417-423
    _chk(oracle.ranlux(100)[:5], [0.94589490, 0.47347850, 0.95152789, 0.42971975, 0.09127384])
    _chk(oracle.ranlux(100)[:5], [0.02618265, 0.03775346, 0.97274780, 0.13302165, 0.43126065])


def test_luxury_p75_and_continuation(oracle):
    oracle.rluxgo(75, 0)  # ranlux.f:425-431
    _chk(oracle.ranlux(100)[:5], [0.53981817, 0.76155043, 0.06029940, 0.79600263, 0.30631220])
    _chk(oracle.ranlux(100)[:5], [0.25600731, 0.23443210, 0.59164381, 0.59035838, 0.07011414])
    # ranlux.f:433-445: the stream simply continues after the status dump
    _chk(oracle.ranlux(100)[:5], [0.22617835, 0.60655993, 0.86417443, 0.43920082, 0.23382509])
    _chk(oracle.ranlux(100)[:5], [0.08107197, 0.21466845, 0.84856731, 0.94078046, 0.85626233])


def test_restart_by_skipping(oracle):
    oracle.rluxgo(4, 7674985)  # ranlux.f:459-467
    for _ in range(10):
        oracle.ranlux(1000)
    kount = oracle.lib().fmc_oracle_rluxat_kount()
    assert kount == 161840
    a = oracle.ranlux(200)
    assert abs(a[0] - 0.019648) < 5.1e-7 and abs(a[199] - 0.590586) < 5.1e-7
    oracle.rluxgo(4, 7674985, kount, 0)
    b = oracle.ranlux(200)
    assert np.array_equal(a, b)


def test_uniforms_are_24_bit_or_padded(oracle):
    oracle.initialise_rng()  # utils.f90:338-342: luxury 3, seed 31415
    u = oracle.ranlux(20000)
    assert (u > 0).all() and (u < 1).all()
    big = u[u >= 2.0 ** -12]
    assert np.array_equal(big * 2 ** 24, np.round(big * 2 ** 24))  # ranlux.f:127-133


def test_rang_polar_method(oracle):
    # utils.f90:345-364: pairs of uniforms, v=2u-1, accept 0<rad<=1, f=sqrt(-2 ln(rad)/rad)
    oracle.initialise_rng()
    z = oracle.rang(7)  # odd n: the second deviate of the last pair is dropped
    oracle.initialise_rng()
    out, k = [], 0
    while len(out) < 8:
        u = oracle.ranlux(2)
        v1, v2 = 2 * u[0] - 1, 2 * u[1] - 1
        rad = v1 * v1 + v2 * v2
        if 0 < rad <= 1:
            f = np.sqrt(-2 * np.log(rad) / rad)
            out += [v1 * f, v2 * f]
    assert np.allclose(z, out[:7], rtol=1e-15)
    oracle.initialise_rng()
    z = oracle.rang(200000)
    assert abs(z.mean()) < 4 / np.sqrt(len(z)) and abs(z.std() - 1) < 4 / np.sqrt(2 * len(z))
    assert abs((z ** 4).mean() - 3) < 0.1
