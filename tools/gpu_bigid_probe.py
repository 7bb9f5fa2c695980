import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import fitter_mc_b200 as fmc
import datasets, oracle_lib as oracle
model, x, y, err = datasets.dataset("c2")
start, _ = oracle.fit_data(model, x, y, err, oracle.model_info(model)["params_init"])
# the start that bench.py derives on the GPU for world > 1
ctx = fmc.Context(0); ctx.set_data(model, x, y, err)
gstart, _ = fmc.fit_data(x, y, err, fmc.initialise_model(model)["params_init"], model)
print("start oracle", start, "gpu", gstart, flush=True)
world = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for step in [1000, 2000, 2001, 2002]:
    for r in range(world):
        first = (step * world + r) * 1_000_000
        ctx.set_data(model, x, y, err)
        ctx.launch(gstart, 1_000_000, first_counter=first); out = ctx.finish()
        print(step, r, first, "ms %.1f" % ctx.kernel_ms(), "passes/fit %.3f" % (out["totals"][3] / 1e6), out["rc_count"][3:11], flush=True)
