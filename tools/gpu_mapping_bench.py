"""Fit mapping by measurement (BASELINE.json north star: "one fit per thread or per warp, chosen by
measurement"): fits/s of one fit per THREAD vs one fit per WARP on each config, one GPU, device
resident, plus the algorithmic roofline fraction of each.
Run on a GPU box:  python tools/gpu_mapping_bench.py [config ...] > gpurun_out/mapping.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datasets  # noqa: E402
import fitter_mc_b200 as fmc  # noqa: E402
import oracle_lib as oracle  # noqa: E402

# flops of one model_y call, transcendentals = 20
C_M = {0: 23, 1: 25, 2: 3 * (20 + 6) + 1, 3: 2 * 20 + 5}
NRES = {"c1": 4_000_000, "c2p3": 2_000_000, "c2": 2_000_000, "c5mild": 2_000_000, "c5": 400_000, "c4": 200_000}


def w_fit(model, n, p, F, G, A):
    cm = C_M[model]
    return n * ((F + p * G) * (cm + 2) + G * (4 * p + p * (p + 1)) + 2 * F + 2 * p * A)


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or ["c4", "c2", "c5", "c1"]
    scale = float(os.environ.get("FMC_BENCH_SCALE", "1"))
    peak, _ = fmc.measure_dfma_peak(0)
    out = {"dfma_peak_tflops": peak}
    for name in names:
        model, x, y, err = datasets.dataset(name)
        info = fmc.initialise_model(model)
        if name == "c5":   # committed start: the initial fit of the stress data is chaotic (tests/datasets.py)
            start = datasets.C5_FITTED.copy()
        else:
            start, _ = oracle.fit_data(model, x, y, err, info["params_init"])
        nres = int(NRES[name] * scale)
        res = {}
        for mapping in ("thread", "warp"):
            ctx = fmc.Context(0)
            ctx.set_data(model, x, y, err)
            ctx.set_fit_mapping(mapping)
            best = None
            for rep in range(3):
                ctx.launch(start, nres, first_counter=rep * nres)
                r = ctx.finish()
                ms = ctx.kernel_ms()
                best = ms if best is None else min(best, ms)
            tot = r["totals"].astype(float) / nres   # nfcall, ngcall, niter, passes per fit
            F, G = tot[0], tot[1]
            w = w_fit(model, len(x), info["nparam"], F, G, G - 2)
            fps = nres / (best * 1e-3)
            res[mapping] = dict(fits_per_s=fps, ms=best, nresample=nres, passes_per_fit=tot[3], F=F, G=G,
                                w_fit_mflop=w / 1e6, algorithmic_tflops=w * fps / 1e12,
                                roofline_frac=w * fps / 1e12 / peak, geometry=ctx.geometry(),
                                rc_count=r["rc_count"][3:11].tolist(), anomalies=int(ctx.anomalies()[2]))
            ctx.close()
        res["warp_over_thread"] = res["warp"]["fits_per_s"] / res["thread"]["fits_per_s"]
        out[name] = res
        print(name, json.dumps(res), file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
