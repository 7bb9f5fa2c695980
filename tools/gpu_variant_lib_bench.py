"""Time the C2 workload with alternative builds of libfmc_b200.so (developer experiments):
python tools/gpu_variant_lib_bench.py lib1.so lib2.so ...  (each in a fresh process)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, json
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import numpy as np
import fitter_mc_b200.api as api
api.lib_path = lambda: %(lib)r
import fitter_mc_b200 as fmc, datasets, oracle_lib as oracle
out = {}
for name in %(configs)r:
    model, x, y, err = datasets.dataset(name)
    start = datasets.C5_FITTED.copy() if name == "c5" else oracle.fit_data(model, x, y, err, fmc.initialise_model(model)["params_init"])[0]
    ctx = fmc.Context(0); ctx.set_data(model, x, y, err)
    nres, best = {"c5": 400000, "c4": 100000}.get(name, 2000000), None
    for rep in range(4):
        ctx.launch(start, nres, first_counter=rep * nres); r = ctx.finish()
        ms = ctx.kernel_ms(); best = ms if best is None else min(best, ms)
    out[name] = dict(fits_per_s=nres / (best * 1e-3), geometry=ctx.geometry(), mean0=float(r["sum_p"][0] / nres))
    ctx.close()
print(json.dumps(out))
"""


def main():
    res = {}
    for lib in sys.argv[1:]:
        configs = tuple(os.environ.get("FMC_VARIANT_CONFIGS", "c2").split(","))
        p = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT, lib=os.path.abspath(lib), configs=configs)],
                           capture_output=True, text=True)
        res[os.path.basename(lib)] = json.loads(p.stdout.strip().splitlines()[-1]) if p.returncode == 0 else p.stderr[-800:]
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
