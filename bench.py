#!/usr/bin/env python
"""bench.py -- bootstrap fits/sec of the B200 engine on BASELINE.json's headline config.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

A "step" is one pass of the hot path (bootstrap.f90:297-330) over one batch of RESAMPLES
resamples per GPU of the C2 workload (model test4: y=a+b*exp(-al*x)+c*x, p=4, ndata=1000;
BASELINE.json configs[1]).  Scaling is weak: every rank takes a disjoint range of global
resample ids of the same length and the ranks' moment accumulators are combined with ONE
NCCL all-reduce per step.

`value`  : whole-job fits/s with the data resident on the GPU, timed with CUDA events on the
           launching stream (fit kernel + all-reduce), max over ranks.
`e2e`    : the same metric through the public host API with HOST buffers: per step the data
           (x, y, err, params_start) go host->device and the accumulator block comes back.
`roofline`: FP64-FMA bound (no dense contraction, working set in shared memory).
           achieved = algorithmic flops per fit (SURVEY.md 8d formula, with the solver's own
           evaluation counters measured in this run) x fits/s; peak = DFMA-chain microbenchmark
           measured in this run (MEASURED_PEAKS.json has no FP64 entry).
`cpu_baseline`: the CPU oracle (C++ restatement of the reference algorithm, reference-equivalent
           cost: serial RANLUX under a lock, covariance dead work, Jacobian copies, histogram
           writes) on the box's host cores, bounded sample.  kind = "port": the reference is
           Fortran and cannot be compiled in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

MODEL = 1            # test4
NDATA = 1000
RESAMPLES = 1_000_000  # per GPU per step (BASELINE.json configs[1])
SEED = 31415
WORKLOAD = "C2: model test4 y=a+b*exp(-al*x)+c*x (p=4, nprop=1), ndata=1000, 1e6 resamples per GPU per step"
C_T = 20  # flops per transcendental, fixed by SURVEY.md section 8d
RESULT_OUT = sys.stdout


def make_problem():
    import datasets

    x, y, err = datasets.c2(model=MODEL, n=NDATA)
    return x, y, err


def algorithmic_flops_per_fit(n, p, F, G, A, c_m):
    """W_fit = n[(F+pG)(c_m+2) + G(4p+p(p+1)) + 2F + 2pA]."""
    return n * ((F + p * G) * (c_m + 2) + G * (4 * p + p * (p + 1)) + 2 * F + 2 * p * A)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[6]))
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


def run_reference(args, rank, world):
    """The reference algorithm on host cores (oracle port; gfortran is not available here)."""
    if rank != 0:
        return
    import oracle_lib as oracle

    x, y, err = make_problem()
    info = oracle.model_info(MODEL)
    start, _ = oracle.fit_data(MODEL, x, y, err, info["params_init"])
    cores = os.cpu_count() or 1
    flags = oracle.COVARIANCE | oracle.COPY_JAC | oracle.HISTOGRAM  # reference-equivalent cost
    sample = int(os.environ.get("FMC_REF_SAMPLE", max(2000, 2500 * cores)))
    oracle.initialise_rng()
    for _ in range(args.warmup):
        oracle.bootstrap(MODEL, x, y, err, start, max(sample // 10, 200), flags=flags, nthreads=cores,
                         per_resample=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.bootstrap(MODEL, x, y, err, start, sample, flags=flags, nthreads=cores, per_resample=False)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = "%d resamples/step of the C2 workload, serial RANLUX+polar stream, covariance+Jacobian-copy+histogram cost kept" % sample
    RESULT_OUT.write(json.dumps({
        "impl": "reference", "metric": "bootstrap fits/sec (ndata=1000, 4 params)", "value": value, "unit": "fits/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": desc},
        "cpu_baseline": {"value": value, "unit": "fits/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }) + "\n")
    RESULT_OUT.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--resamples", type=int, default=RESAMPLES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # The contract is ONE JSON line on stdout.  Libraries chat on fd 1 (NCCL prints its version
    # banner there), so keep the real stdout aside for the result line and send the rest to stderr.
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if os.environ.get("FMC_BENCH_WATCHDOG"):  # debugging aid: dump every thread's stack if a run stalls
        import faulthandler

        faulthandler.dump_traceback_later(int(os.environ["FMC_BENCH_WATCHDOG"]), repeat=False, exit=True)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import fitter_mc_b200 as fmc
    from fitter_mc_b200 import distributed as fdist
    import oracle_lib as oracle  # cpu_baseline leg only

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: fitter_mc_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    x, y, err = make_problem()
    info = fmc.initialise_model(MODEL)
    p, q = info["nparam"], info["nprop"]
    start, _ = fmc.fit_data(x, y, err, info["params_init"], MODEL) if local_rank == 0 and world == 1 else (None, None)
    ctx = fmc.Context(local_rank)
    ctx.set_data(MODEL, x, y, err)
    if start is None:  # every rank derives the same start from the same data with its own GPU
        zeros = torch.zeros(NDATA, dtype=torch.float64, device=dev)  # offset_data(.FALSE.)
        torch.cuda.synchronize()
        ctx.launch(info["params_init"], 1, z_inject_dev=zeros.data_ptr(), keep=True)
        start = ctx.finish()["params"][0]
    nres = args.resamples
    # a dedicated (non-default) torch stream carries the fit kernel, the all-reduce and the
    # timing events, so the CUDA events see exactly the work they bracket
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    acc = torch.zeros(ctx.accumulator_len(), dtype=torch.float64, device=dev)
    ctx.set_accumulator_dev(acc.data_ptr())
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def device_step(step_no):
        """kernel + all-reduce with the data resident; returns the step's device time (ms)."""
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.launch(start, nres, seed=SEED, first_counter=(step_no * world + rank) * nres)
        fdist.allreduce_accumulator(acc)  # the single collective of the path: 31 doubles over NVLink
        e1.record(stream)
        e1.synchronize()
        ctx.wait()
        return e0.elapsed_time(e1)

    sync_token = torch.zeros(1, dtype=torch.float64, device=dev)

    def sync_all():
        """barrier + device synchronize (the barrier is a 1-element all-reduce on the working stream)"""
        if world > 1:
            dist.all_reduce(sync_token)
        torch.cuda.synchronize()

    for w in range(args.warmup):
        device_step(w)

    sampler = ClockSampler(local_rank)
    sync_all()
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    ms = [device_step(args.warmup + k) for k in range(args.steps)]
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = float(sum(ms))
    kernel_ms = ctx.kernel_ms()
    last = fmc.decode_accumulator(MODEL, start, acc.cpu().numpy())
    n_anomalies = ctx.anomalies()[2]  # fits of the last step that hit a solver safety cap (normally 0)

    # ---- e2e: public host API with HOST buffers; per step the inputs go host->device and the
    # (all-reduced) accumulator block comes back device->host, all inside the timed region and
    # all on the one stream that also carries the kernels and the collective.
    def e2e_step(step_no):
        ctx.set_data(MODEL, x, y, err)              # H2D of the step's inputs (x, y, err)
        ctx.launch(start, nres, seed=SEED, first_counter=(step_no * world + rank) * nres)  # H2D of params_start
        fdist.allreduce_accumulator(acc)
        host_block = acc.cpu()                      # D2H of the result block (synchronises)
        ctx.wait()
        return host_block

    e2e_step(1000)
    sync_all()
    t0 = time.perf_counter()
    for k in range(args.steps):
        last_block = e2e_step(2000 + k)
    sync_all()
    e2e_s = time.perf_counter() - t0
    e2e_check = fmc.decode_accumulator(MODEL, start, last_block.numpy())
    assert e2e_check["nfit"] == nres * world, (e2e_check["nfit"], nres, world)

    # max over ranks
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, t_wall], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, t_wall = [float(v) for v in t.cpu()]
    total_fits = float(nres) * args.steps * world
    value = total_fits / (dev_ms * 1e-3)
    e2e_value = total_fits / e2e_s
    if world > 1:
        sync_all()
        dist.destroy_process_group()

    if rank == 0:
        # FP64 roofline denominator: DFMA-chain microbenchmark on this GPU, after the timed work
        peak_tflops, _ = fmc.measure_dfma_peak(local_rank, 20000)
        tot = last["totals"].astype(float) / max(last["nfit"], 1)
        F, G = tot[0], tot[1]
        A = max(G - 2.0, 0.0)
        c_m = C_T + 5  # test4: one exp + 5 flops
        w_fit = algorithmic_flops_per_fit(NDATA, p, F, G, A, c_m)
        achieved = w_fit * (nres / (kernel_ms * 1e-3)) / 1e12
        line = {
            "metric": "bootstrap fits/sec (ndata=1000, 4 params)", "value": value, "unit": "fits/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "resamples_per_gpu_per_step": nres, "seed": SEED,
                       "jacobian": "forward differences (nl2sno, the reference's path)",
                       "l2": "flushed between steps (256 MiB memset); working set (32 KB) lives in shared memory",
                       "geometry": ctx.geometry(), "per_fit": {"nfcall": F, "ngcall": G, "niter": tot[2], "passes": tot[3]},
                       "wall_s_timed_region": t_wall, "solver_safety_cap_hits_last_step": int(n_anomalies)},
            "e2e": {"value": e2e_value, "unit": "fits/s", "h2d_bytes_per_step": int(8 * (3 * NDATA + p)),
                    "d2h_bytes_per_step": int(8 * ctx.accumulator_len())},
            "gpu_launches": 5 * args.steps * world,  # jac0, init_shared, iterate, init_pass, iterate (fit_data = two nl2sno calls)
            "clocks": clocks,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / peak_tflops,
                         # dram__bytes_read+write of the dominant kernel (iterate, first nl2sno call) per launch,
                         # from profiles/r1_pipeline_ncu_summary.txt; algorithmic bytes are ~0.4 KB of pipeline
                         # state per fit (0.4 GB per launch), the rest is the solver's thread-local state
                         "traffic": 2.01e9,
                         "note": "FP64-FMA bound (SURVEY.md 8d), not hbm/tensor; peak = DFMA-chain microbenchmark "
                                 "measured in this run; achieved = %.0f algorithmic flops/fit x kernel fits/s" % w_fit},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = int(os.environ.get("FMC_REF_SAMPLE", max(2000, 2500 * cores)))
            flags = oracle.COVARIANCE | oracle.COPY_JAC | oracle.HISTOGRAM
            oracle.initialise_rng()
            t0 = time.perf_counter()
            oracle.bootstrap(MODEL, x, y, err, start, sample, flags=flags, nthreads=cores, per_resample=False)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample / dt, "unit": "fits/s", "cores": cores, "kind": "port",
                                    "sample": "%d resamples of the C2 workload, reference-equivalent cost "
                                              "(serial RANLUX, covariance, Jacobian copies, histogram)" % sample}
        RESULT_OUT.write(json.dumps(line) + "\n")
        RESULT_OUT.flush()


if __name__ == "__main__":
    main()
