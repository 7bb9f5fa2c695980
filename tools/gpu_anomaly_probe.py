import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import fitter_mc_b200 as fmc
import datasets
model, x, y, err = datasets.dataset("c2")
ctx = fmc.Context(0); ctx.set_data(model, x, y, err)
gstart, _ = fmc.fit_data(x, y, err, fmc.initialise_model(model)["params_init"], model)
found = []
ranges = [8_004_000_000 + i * 1_000_000 for i in range(8)] + [16_000_000_000 + i * 1_000_000 for i in range(24)]
for first in ranges:
    ctx.launch(gstart, 1_000_000, first_counter=first); out = ctx.finish()
    ids, flags, n = ctx.anomalies()
    print(first, "ms %.1f" % ctx.kernel_ms(), "anomalies", n, list(zip(ids.tolist(), flags.tolist())), flush=True)
    found += list(zip(ids.tolist(), flags.tolist()))
np.save("gpurun_out/anom_start.npy", gstart)
if found:
    zs = np.array([fmc.generate_offsets(fmc.DEFAULT_SEED, i, 1, len(x))[0] for i, _ in found[:16]])
    np.save("gpurun_out/anom_z.npy", zs)
    np.save("gpurun_out/anom_ids.npy", np.array(found[:16]))
    for (i, f), z in zip(found[:4], zs):
        r = fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=1, first_counter=i, params_init=gstart,
                                     initial_fit=False, per_resample=True)
        print("replay", i, f, r["params"], r["rc_count"][3:11], r["totals"], flush=True)
