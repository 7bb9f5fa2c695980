"""Schedule by measurement: fits/s of the refill kernel vs the round schedule (csrc/round_kernels.cuh) on
each config, one GPU, device resident, with the algorithmic roofline fraction.
Run on a GPU box:  python tools/gpu_schedule_bench.py [config[:nres] ...] > gpurun_out/schedule.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datasets  # noqa: E402
import fitter_mc_b200 as fmc  # noqa: E402
import oracle_lib as oracle  # noqa: E402

C_M = {0: 23, 1: 25, 2: 3 * (20 + 6) + 1, 3: 2 * 20 + 5}
NRES = {"c1": 4_000_000, "c2p3": 2_000_000, "c2": 2_000_000, "c5mild": 2_000_000, "c5": 1_000_000, "c4": 200_000}


def w_fit(model, n, p, F, G, A):
    cm = C_M[model]
    return n * ((F + p * G) * (cm + 2) + G * (4 * p + p * (p + 1)) + 2 * F + 2 * p * A)


def main():
    specs = [a for a in sys.argv[1:] if not a.startswith("-")] or ["c2", "c5", "c4", "c1"]
    peak, _ = fmc.measure_dfma_peak(0)
    out = {"dfma_peak_tflops": peak, "round_waves": os.environ.get("FMC_ROUND_WAVES", "8")}
    for spec in specs:
        name, _, nr = spec.partition(":")
        model, x, y, err = datasets.dataset(name)
        info = fmc.initialise_model(model)
        start = datasets.C5_FITTED.copy() if name == "c5" else oracle.fit_data(model, x, y, err, info["params_init"])[0]
        nres = int(nr) if nr else NRES[name]
        res = {}
        for sched in ("refill", "rounds"):
            ctx = fmc.Context(0)
            ctx.set_data(model, x, y, err)
            ctx.set_schedule(sched)
            best = None
            for rep in range(3):
                ctx.launch(start, nres, first_counter=rep * nres)
                r = ctx.finish()
                ms = ctx.kernel_ms()
                best = ms if best is None else min(best, ms)
            tot = r["totals"].astype(float) / nres
            F, G = tot[0], tot[1]
            w = w_fit(model, len(x), info["nparam"], F, G, G - 2)
            fps = nres / (best * 1e-3)
            res[sched] = dict(fits_per_s=fps, ms=best, nresample=nres, passes_per_fit=tot[3], F=F, G=G,
                              algorithmic_tflops=w * fps / 1e12, roofline_frac=w * fps / 1e12 / peak,
                              geometry=ctx.geometry(), launches=ctx.launch_count(), mean0=float(r["sum_p"][0] / nres),
                              rc_count=r["rc_count"][3:11].tolist(), anomalies=int(ctx.anomalies()[2]))
            ctx.close()
        res["rounds_over_refill"] = res["rounds"]["fits_per_s"] / res["refill"]["fits_per_s"]
        out[name] = res
        print(name, json.dumps(res), file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
