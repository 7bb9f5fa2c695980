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


def c5(n=1000, rel=0.05, floor=0.01):
    return _make(3, np.logspace(-1, np.log10(50), n), lambda y: rel * np.abs(y) + floor, 5000)


def dataset(name):
    return {"c1": (0,) + c1_probe(), "c2p3": (0,) + c2(0), "c2": (1,) + c2(1), "c4": (2,) + c4(),
            "c5": (3,) + c5()}[name]
