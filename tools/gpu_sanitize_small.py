"""Tiny end-to-end run for compute-sanitizer (memcheck): every kernel of the pipeline, generator
and injection paths, per-resample outputs, ragged n, two chunks' worth of code paths."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import fitter_mc_b200 as fmc
import datasets

model, x, y, err = datasets.dataset("c1")
x, y, err = x[:37], y[:37], err[:37]
start, rc = fmc.fit_data(x, y, err, [0.0, 2.0, 1.0], model)
a = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=700, params_init=start, initial_fit=False, per_resample=True)
z = fmc.generate_offsets(31415, 0, 700, len(x))
b = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=700, params_init=start, initial_fit=False, per_resample=True,
                             z_inject=z)
assert np.array_equal(a["params"], b["params"])
m3, x3, y3, e3 = datasets.dataset("c5mild")
c = fmc.monte_carlo_fit_data(x3[:200], y3[:200], e3[:200], m3, nrand_points=300, per_resample=True)
print("sanitize run ok", a["params_opt"], c["params_opt"])
