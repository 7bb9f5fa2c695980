"""Throughput of the optional modes on the C2 workload, one GPU: forward differences (default),
analytic Jacobian (row f3), device-side binning (row f4), per-resample outputs kept.
Run on a GPU box: python tools/gpu_modes_bench.py > gpurun_out/modes.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datasets  # noqa: E402
import fitter_mc_b200 as fmc  # noqa: E402
import oracle_lib as oracle  # noqa: E402


def main():
    out = {}
    for name in ("c2", "c1", "c5mild"):
        model, x, y, err = datasets.dataset(name)
        info = fmc.initialise_model(model)
        start, _ = oracle.fit_data(model, x, y, err, info["params_init"])
        nres = 2_000_000
        ctx = fmc.Context(0)
        ctx.set_data(model, x, y, err)
        ctx.launch(start, 4096, keep=True)
        pilot = ctx.finish()
        lo, hi = fmc.histogram_range_from_pilot(np.column_stack([pilot["params"], pilot["props"]]))
        res = {}
        for label, jac, bins, keep in (("forward_difference", 0, 0, False), ("analytic", 1, 0, False),
                                       ("forward_difference+binned_1024", 0, 1024, False),
                                       ("forward_difference+per_resample_outputs", 0, 0, True)):
            ctx.set_jacobian_mode(jac)
            ctx.set_histogram(bins, lo, hi)
            best = None
            for rep in range(3):
                ctx.launch(start, nres, first_counter=rep * nres, keep=keep)
                r = ctx.finish()
                ms = ctx.kernel_ms()
                best = ms if best is None else min(best, ms)
            res[label] = dict(fits_per_s=nres / (best * 1e-3), ms=best, passes_per_fit=float(r["totals"][3]) / nres)
            if bins:
                h = ctx.histogram()
                res[label]["outside_range"] = int(h[:, 0].sum() + h[:, -1].sum())
        ctx.close()
        out[name] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
