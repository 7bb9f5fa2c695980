"""Large runs on one GPU: 1e8 resamples of C2 (chunked pipeline, compensated sums), and the
C4 / C5 configs at sizes that fill the device.  Prints fits/s and ensemble statistics."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import fitter_mc_b200 as fmc
import datasets, oracle_lib as oracle

out = {}
for name, nres in [("c2", 100_000_000), ("c2p3", 10_000_000), ("c5", 10_000_000), ("c4", 200_000), ("c1", 100_000)]:
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    model, x, y, err = datasets.dataset(name)
    info = oracle.model_info(model)
    start = datasets.C5_FITTED.copy() if name == "c5" else oracle.fit_data(model, x, y, err, info["params_init"])[0]
    ctx = fmc.Context(0)
    ctx.set_data(model, x, y, err)
    t0 = time.perf_counter()
    ctx.launch(start, nres)
    res = ctx.finish()
    wall = time.perf_counter() - t0
    ms = ctx.kernel_ms()
    mean, sd = fmc.finalise_moments(nres, res["sum_p"], res["sum_p2"])
    qm, qs = fmc.finalise_moments(nres, res["sum_prop"], res["sum_prop2"])
    rec = dict(config=name, n=len(x), p=info["nparam"], nresample=nres, kernel_ms=ms, wall_s=wall,
               fits_per_s=nres / ms * 1e3, mean=mean.tolist(), sd=sd.tolist(), prop_mean=qm.tolist(), prop_sd=qs.tolist(),
               rc_count=res["rc_count"][3:11].tolist(), per_fit=(res["totals"] / nres).tolist(), fitted=start.tolist(),
               schedule=ctx.schedule(), safety_cap_hits=int(ctx.anomalies()[2]))
    out[name] = rec
    print(json.dumps(rec))
    ctx.close()
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "large_runs.json"), "w"), indent=1)
