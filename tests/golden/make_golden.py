"""Generates the golden fixtures of tests/golden/ from the reference's own Python implementation
(/root/reference/mcfit.py), the only reference-shipped code that runs in this image.

    python tests/golden/make_golden.py        (needs /root/reference; run in the build container)

1. mcfit_probe_stdout.txt  -- stdout of `mcfit.py probe.dat` on the survey's probe file
   (BASELINE.md section 2).  The "optimised parameters", chi^2 and Q lines are deterministic; the
   Monte-Carlo block is not (mcfit.py draws unseeded numpy normals) and is kept for reference only.
2. mcfit_curve_fit.json    -- the reference's per-resample fit (mcfit.py:96-99:
   scipy.optimize.curve_fit(model_y, xx, yy + z*err, sigma=err, p0=initparams)) on 40 SEEDED
   resamples of the probe data, so that tests can feed the same resamples to the oracle.
3. ranlux_kat.json         -- the expected LUXTST output from the comments of ranlux.f:399-473.
"""
import importlib.util
import io
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.dirname(HERE))
import datasets  # noqa: E402

x, y, err = datasets.c1_probe()
with tempfile.TemporaryDirectory() as tmp:
    np.savetxt(os.path.join(tmp, "probe.dat"), np.c_[x, y, err])
    out = subprocess.run([sys.executable, os.path.join(REF, "mcfit.py"), "probe.dat"], cwd=tmp, capture_output=True,
                         text=True, check=True).stdout
open(os.path.join(HERE, "mcfit_probe_stdout.txt"), "w").write(out)

# the reference's model_y (mcfit.py:33-44), taken from its source text so that it is THEIR function
src = open(os.path.join(REF, "mcfit.py")).read()
m = re.search(r"^def model_y\(.*?(?=^def |\Z)", src, flags=re.S | re.M)
ns = {"np": np}
exec(m.group(0), ns)
model_y = ns["model_y"]
from scipy.optimize import curve_fit  # noqa: E402

init, _ = curve_fit(model_y, x, y, sigma=err, p0=[0.0, 2.0, 1.0])
rng = np.random.default_rng(20261019)
z = rng.normal(size=(40, len(x)))
fits = [curve_fit(model_y, x, y + zi * err, sigma=err, p0=init)[0].tolist() for zi in z]
json.dump({"source": "mcfit.py:96-99 (scipy %s)" % __import__("scipy").__version__, "seed": 20261019,
           "initparams": init.tolist(), "params": fits}, open(os.path.join(HERE, "mcfit_curve_fit.json"), "w"), indent=1)

json.dump({
    "source": "ranlux.f:399-473 (expected output of LUXTST)",
    "default_seed_lux3": {"1-5": [0.53981817, 0.76155043, 0.06029940, 0.79600263, 0.30631220],
                          "101-105": [0.43156743, 0.03774416, 0.24897110, 0.00147784, 0.90274453]},
    "lux0_default_seed": {"1-5": [0.53981817, 0.76155043, 0.06029940, 0.79600263, 0.30631220],
                          "101-105": [0.41538775, 0.05330932, 0.58195311, 0.91397446, 0.67034441]},
    "p389_seed1": {"1-5": [0.94589490, 0.47347850, 0.95152789, 0.42971975, 0.09127384],
                   "101-105": [0.02618265, 0.03775346, 0.97274780, 0.13302165, 0.43126065]},
    "p75_default_seed": {"1-5": [0.53981817, 0.76155043, 0.06029940, 0.79600263, 0.30631220],
                         "101-105": [0.25600731, 0.23443210, 0.59164381, 0.59035838, 0.07011414],
                         "201-205": [0.22617835, 0.60655993, 0.86417443, 0.43920082, 0.23382509],
                         "301-305": [0.08107197, 0.21466845, 0.84856731, 0.94078046, 0.85626233]},
    "lux4_seed7674985": {"kount_after_10000": 161840, "next": 0.019648, "200th": 0.590586},
}, open(os.path.join(HERE, "ranlux_kat.json"), "w"), indent=1)
print("wrote golden fixtures to", HERE)
