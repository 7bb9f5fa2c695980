"""Synthetic inputs for the BASELINE.json configs.

All data come from numpy default_rng(seed) (PCG64), as the survey's recipe says.  Model ids:
0 = reference default (p=3), 1 = 4-parameter extension, 2 = 10-parameter three-peak model,
3 = stiff exponential + power law.
"""
import numpy as np

TRUTH = {
    0: np.array([0.5, 2.0, 1.3]),
    1: np.array([0.5, 2.0, 1.3, 0.1]),
    2: np.array([1.0, 5.0, 2.5, 0.4, 3.0, 5.0, 0.6, 4.0, 7.5, 0.5]),
    3: np.array([5.0, 2.0, 0.5, 1.5]),
}


def model_y(model, p, x):
    x = np.asarray(x, dtype=np.float64)
    if model == 0:
        return p[0] + p[1] * np.exp(-p[2] * x)
    if model == 1:
        return p[0] + p[1] * np.exp(-p[2] * x) + p[3] * x
    if model == 2:
        y = np.full_like(x, p[0])
        for k in range(3):
            u = (x - p[2 + 3 * k]) / p[3 + 3 * k]
            y = y + p[1 + 3 * k] * np.exp(-0.5 * (u * u))
        return y
    if model == 3:
        return p[0] * np.exp(-p[1] * x) + p[2] * np.power(x, -p[3])
    raise ValueError(model)


def _make(model, x, errf, seed):
    rng = np.random.default_rng(seed)
    y0 = model_y(model, TRUTH[model], x)
    err = errf(y0)
    y = y0 + rng.normal(0.0, 1.0, len(x)) * err
    return x, y, err


def c1_probe():
    """C1: the survey's probe file."""
    rng = np.random.default_rng(12345)
    x = np.linspace(0, 5, 40)
    err = np.full(40, 0.05)
    y = 0.5 + 2.0 * np.exp(-1.3 * x) + rng.normal(0, err)
    return x, y, err


def c2(model=1, n=1000):
    """C2/C3: n=1000, x in [0,5], err=0.05, data seed 1000; model 0 (p=3) or 1 (p=4)."""
    return _make(model, np.linspace(0, 5, n), lambda y: np.full_like(y, 0.05), 1000)


def c4(n=10000):
    return _make(2, np.linspace(0, 10, n), lambda y: np.full_like(y, 0.1), 10000)


# C5 ("tune err until >= 10 % of fits end with NL2SOL codes 7-10, measured by the
# oracle").  Measured with the oracle on 2000 resamples (tests/test_c5_stress.py pins it):
#   err = 0.05|y| + 0.01 (the untuned recipe)   0.0 % codes 7-10, every fit 6-8 passes
#   err = 12.5|y| + 1.0                          4.3 %
#   err = 13|y| + 1.0  (C5_REL, C5_FLOOR)        23 %: 22 % singular convergence (7), the rest false
#                                                convergence (8) and the evaluation / iteration limits
#                                                (9, 10); 5 ... 500 passes per fit, mean 77
# At that noise level the exponential term is not identifiable: the fit to the un-resampled data
# itself sits in the degenerate basin (a ~ -8.6e3, b ~ 45: a spike at the first data points) and
# resampled fits run along flat valleys (a down to -1e16) -- the regime in which the reference
# averages whatever nl2sno returns (bootstrap.f90:215-226 never looks at iv(1)).  Parameters of
# such fits are not reproducible between ANY two builds; their chi-squared is
# (tests/test_c5_stress.py).  "c5mild" is the untuned recipe, used where per-fit parameter
# parity is asserted.
C5_REL, C5_FLOOR = 13.0, 1.0
# The fit to the un-resampled C5 data (bootstrap.f90:288-291) as the oracle returns it in the build
# container, from params_init = (1, 1, 1, 1): 42 evaluations into the degenerate basin.  That path is
# chaotic (other libm builds of exp/log send it elsewhere, and the reference's gqtstp can spin forever
# on it -- tests/test_robustness.py), so every test and bench of C5 starts its resamples from these
# committed values instead of refitting.
C5_FITTED = np.array([float.fromhex("-0x1.0cbd778212d33p+13"), float.fromhex("0x1.6c15b0c6968e6p+5"),
                      float.fromhex("0x1.fb0bea4540d5ep-2"), float.fromhex("0x1.c90048cca64bap-1")])


def c5(n=1000, rel=C5_REL, floor=C5_FLOOR):
    return _make(3, np.logspace(-1, np.log10(50), n), lambda y: rel * np.abs(y) + floor, 5000)


def c5_mild(n=1000):
    return c5(n, rel=0.05, floor=0.01)


def dataset(name):
    return {"c1": (0,) + c1_probe(), "c2p3": (0,) + c2(0), "c2": (1,) + c2(1), "c4": (2,) + c4(),
            "c5": (3,) + c5(), "c5mild": (3,) + c5_mild()}[name]
