"""Short run of one config for ncu captures: python tools/gpu_profile_c2.py <config> <nres> <reps> <mode> <mapping>."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import fitter_mc_b200 as fmc
import datasets, oracle_lib as oracle

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
nres = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mode = sys.argv[4] if len(sys.argv) > 4 else "fd"        # "fd", "analytic", "analytic+binned", "fd+binned"
mapping = sys.argv[5] if len(sys.argv) > 5 else "thread"
model, x, y, err = datasets.dataset(name)
info = oracle.model_info(model)
start = datasets.C5_FITTED.copy() if name == "c5" else oracle.fit_data(model, x, y, err, info["params_init"])[0]
ctx = fmc.Context(0)
ctx.set_data(model, x, y, err)
ctx.set_fit_mapping(mapping)
if "analytic" in mode:
    ctx.set_jacobian_mode("analytic")
if "binned" in mode:
    import numpy as np
    ctx.launch(start, 4096, keep=True)
    pilot = ctx.finish()
    lo, hi = fmc.histogram_range_from_pilot(np.column_stack([pilot["params"], pilot["props"]]))
    ctx.set_histogram(1024, lo, hi)
for rep in range(reps):
    ctx.launch(start, nres, first_counter=rep * nres)
    out = ctx.finish()
    print(name, nres, "kernel ms %.3f -> %.4e fits/s" % (ctx.kernel_ms(), nres / ctx.kernel_ms() * 1e3), ctx.geometry())
