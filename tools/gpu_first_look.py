"""Scratch driver for the first GPU runs: DFMA peak, kernel geometry, throughput per model."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import fitter_mc_b200 as fmc
import datasets, oracle_lib as oracle

print("dfma peak TFLOP/s:", [fmc.measure_dfma_peak(0, 20000) for _ in range(3)])
for name, nres in [("c1", 1 << 20), ("c2p3", 1 << 19), ("c2", 1 << 19), ("c5", 1 << 17), ("c4", 1 << 13)]:
    model, x, y, err = datasets.dataset(name)
    info = oracle.model_info(model)
    start, _ = oracle.fit_data(model, x, y, err, info["params_init"])
    ctx = fmc.Context(0)
    ctx.set_data(model, x, y, err)
    for rep in range(3):
        ctx.launch(start, nres, first_counter=rep * nres)
        out = ctx.finish()
        ms = ctx.kernel_ms()
    tot = out["totals"]
    print(name, "n=%d p=%d" % (len(x), info["nparam"]), ctx.geometry(), "kernel ms %.2f -> %.3e fits/s" % (ms, nres / ms * 1e3),
          "per fit: nfcall %.2f ngcall %.2f niter %.2f passes %.2f" % tuple(tot / nres), "rc", out["rc_count"][3:11])
    m, s = fmc.finalise_moments(nres, out["sum_p"], out["sum_p2"])
    print("   mean", m, "sd", s)
    ctx.close()
