#!/usr/bin/env python
"""bench.py -- bootstrap fits/sec of the B200 engine on BASELINE.json's configs.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, headline config C2
    python bench.py --workload c3|c4|c5 ...                  # the other BASELINE.json configs
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

A "step" is one pass of the hot path (bootstrap.f90:297-330) over one batch of resamples:
  c2 (default; BASELINE.json configs[1], the config the metric is quoted on): model test4
      y=a+b*exp(-al*x)+c*x (p=4), ndata=1000, 1e6 resamples per GPU per step  -- weak scaling;
  c3 (configs[2], the north-star target run): the same fit, 1e8 resamples per step split contiguously
      over the N GPUs -- strong scaling; checks the ensemble against the reference's RANLUX stream and
      the N-GPU sums against a 1-GPU run of the same ids;
  c4 (configs[3]): 10-parameter three-peak model, ndata=10^4, 1e7 resamples per GPU per step;
  c5 (configs[4]): stiff exponential + power law with >= 10 % non-converging fits, 1e7 per GPU per step.
Every rank takes a disjoint range of global resample ids; the ranks' moment accumulators are
combined by ONE ncclAllReduce per step, issued by the library itself (fmc_allreduce, C ABI) on the
stream that carries the kernels.  torch.distributed only carries the rendezvous and the barrier.

`value`  : whole-job fits/s with the data resident on the GPU, timed with CUDA events on the
           launching stream (fit kernels + all-reduce), max over ranks.
`e2e`    : the same metric through the public host API with HOST buffers: per step the data
           (x, y, err, params_start) go host->device and the accumulator block comes back.
`roofline`: FP64-FMA bound (no dense contraction, working set in shared memory).
           achieved = algorithmic flops per fit (SURVEY.md 8d formula, with the solver's own
           evaluation counters measured in this run) x fits/s; peak = DFMA-chain microbenchmark
           measured in this run (MEASURED_PEAKS.json has no FP64 entry); hw_counted_tflops and
           traffic come from the committed ncu capture named beside them (profiles/).
`cpu_baseline`: the CPU oracle (C++ restatement of the reference algorithm, reference-equivalent
           cost: serial RANLUX under a lock, covariance dead work, Jacobian copies, histogram
           writes) on the box's host cores, bounded sample.  kind = "port": the reference is
           Fortran and cannot be compiled in this image.  `cpu_baseline_lean` is the same without
           the dead work, `cpu_solver_only` the same without the serial generator, `mcfit_py` the loop
           of the reference's mcfit.py (scipy curve_fit per resample) restated here.
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

SEED = 31415
C_T = 20  # flops per transcendental, fixed by SURVEY.md section 8d
RESULT_OUT = sys.stdout

# c_m = flops of one model_y call with transcendentals counted as C_T
WORKLOADS = {
    "c2": dict(dataset="c2", c_m=C_T + 5, resamples=1_000_000, scaling="weak",
               metric="bootstrap fits/sec (ndata=1000, 4 params)",
               text="C2: model test4 y=a+b*exp(-al*x)+c*x (p=4, nprop=1), ndata=1000, 1e6 resamples per GPU per step"),
    "c3": dict(dataset="c2", c_m=C_T + 5, total=100_000_000, scaling="strong",
               metric="bootstrap fits/sec (ndata=1000, 4 params)",
               text="C3: model test4 (p=4, nprop=1), ndata=1000, 1e8 resamples per step split contiguously over the GPUs"),
    "c4": dict(dataset="c4", c_m=3 * (C_T + 6) + 1, resamples=10_000_000, scaling="weak",
               metric="bootstrap fits/sec (ndata=10000, 10 params)",
               text="C4: constant + three Gaussian peaks (p=10, nprop=3), ndata=10000, 1e7 resamples per GPU per step"),
    "c5": dict(dataset="c5", c_m=2 * C_T + 5, resamples=10_000_000, scaling="weak",
               metric="bootstrap fits/sec (ndata=1000, 4 params, stiff model, >=10% non-converging fits)",
               text="C5: stiff a*exp(-b*x)+c*x^(-d) (p=4, nprop=2), ndata=1000, err tuned to >=10% NL2SOL codes 7-10, "
                    "1e7 resamples per GPU per step"),
}


def make_problem(workload):
    import datasets

    return datasets.dataset(WORKLOADS[workload]["dataset"])


def algorithmic_flops_per_fit(n, p, F, G, A, c_m):
    """W_fit = n[(F+pG)(c_m+2) + G(4p+p(p+1)) + 2F + 2pA]."""
    return n * ((F + p * G) * (c_m + 2) + G * (4 * p + p * (p + 1)) + 2 * F + 2 * p * A)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t_begin = self.t_end = None  # the timed region (perf_counter): only rows that arrive inside it count

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.t_begin = time.perf_counter()

    def end(self):
        self.t_end = time.perf_counter()

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
        inside = [r for t, r in self.rows if self.t_begin is None or (self.t_begin <= t <= (self.t_end or t) + 0.05)]
        if not inside:  # a timed region shorter than the sampling period: the nearest sample
            inside = [r for t, r in self.rows[-1:]]
        for r in inside:
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


def cpu_legs(workload, model, x, y, err, start, cores, which=("reference",)):
    """The reference algorithm on the host cores (oracle port), bounded samples; fits/s per leg."""
    import oracle_lib as oracle

    heavy = WORKLOADS[workload]["dataset"] == "c4"
    sample = int(os.environ.get("FMC_REF_SAMPLE", max(2000, 2500 * cores) // (40 if heavy else 1)))
    out = {}
    for leg in which:
        if leg == "reference":   # reference-equivalent cost
            flags, z, desc = oracle.COVARIANCE | oracle.COPY_JAC | oracle.HISTOGRAM, None, \
                "serial RANLUX+polar stream, covariance + Jacobian-copy + histogram cost kept"
        elif leg == "lean":      # the same without the work the reference computes and never reads
            flags, z, desc = 0, None, "serial RANLUX+polar stream; covariance, Jacobian copies and histogram off"
        else:                    # the solver alone: offsets pre-generated, no serial generator
            flags, desc = 0, "offsets pre-generated (no serial generator), covariance/copies/histogram off"
            z = np.random.default_rng(7).normal(size=(sample, len(x)))
        if WORKLOADS[workload]["dataset"] == "c5":
            flags |= oracle.SAFETY_CAP  # the reference's gqtstp retry loop can spin forever on this data
        oracle.initialise_rng()
        t0 = time.perf_counter()
        oracle.bootstrap(model, x, y, err, start, sample, z=z, flags=flags, nthreads=cores, per_resample=False)
        dt = time.perf_counter() - t0
        out[leg] = {"value": sample / dt, "unit": "fits/s", "cores": cores, "kind": "port",
                    "sample": "%d resamples of the %s workload, %s" % (sample, workload.upper(), desc)}
    return out


def mcfit_py_leg():
    """The loop of the reference's mcfit.py (mcfit.py:90-115: numpy.random.normal + scipy curve_fit per
    resample, nsamples = 2000, mcfit.py:28) restated on the C1 probe file -- the reference file itself is
    not present on the GPU box.  One core."""
    try:
        import scipy.optimize

        import datasets

        x, y, err = datasets.c1_probe()

        def model_y(xx, a, b, alpha):
            return a + b * np.exp(-alpha * xx)

        p0 = [0.0, 2.0, 1.0]
        popt, _ = scipy.optimize.curve_fit(model_y, x, y, sigma=err, p0=p0)
        nsamples = 2000
        rng = np.random.default_rng(0)
        s1, s2 = np.zeros(3), np.zeros(3)
        t0 = time.perf_counter()
        for _ in range(nsamples):
            yp = rng.normal(y, err)
            p, _ = scipy.optimize.curve_fit(model_y, x, yp, sigma=err, p0=popt)
            s1 += p
            s2 += p * p
        dt = time.perf_counter() - t0
        return {"value": nsamples / dt, "unit": "fits/s", "cores": 1, "kind": "restated loop of mcfit.py:90-115",
                "sample": "2000 resamples of the C1 probe file (n=40, p=3), scipy.optimize.curve_fit"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)[:200]}


def run_reference(args, rank, world):
    """The reference algorithm on host cores (oracle port; gfortran is not available here)."""
    if rank != 0:
        return
    import oracle_lib as oracle

    wl = WORKLOADS[args.workload]
    model, x, y, err = make_problem(args.workload)
    info = oracle.model_info(model)
    if wl["dataset"] == "c5":
        import datasets

        start = datasets.C5_FITTED.copy()
    else:
        start, _ = oracle.fit_data(model, x, y, err, info["params_init"])
    cores = os.cpu_count() or 1
    heavy = wl["dataset"] == "c4"
    flags = oracle.COVARIANCE | oracle.COPY_JAC | oracle.HISTOGRAM  # reference-equivalent cost
    if wl["dataset"] == "c5":
        flags |= oracle.SAFETY_CAP  # the reference's gqtstp retry loop can spin forever on this data
    sample = int(os.environ.get("FMC_REF_SAMPLE", max(2000, 2500 * cores) // (40 if heavy else 1)))
    oracle.initialise_rng()
    for _ in range(args.warmup):
        oracle.bootstrap(model, x, y, err, start, max(sample // 10, 50), flags=flags, nthreads=cores,
                         per_resample=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.bootstrap(model, x, y, err, start, sample, flags=flags, nthreads=cores, per_resample=False)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = ("%d resamples per step of the %s workload (a rate: the CPU arm cannot run the full step in minutes), "
            "serial RANLUX+polar stream, covariance+Jacobian-copy+histogram cost kept" % (sample, args.workload.upper()))
    RESULT_OUT.write(json.dumps({
        "impl": "reference", "metric": wl["metric"], "value": value, "unit": "fits/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["text"]},
        "cpu_baseline": {"value": value, "unit": "fits/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }) + "\n")
    RESULT_OUT.flush()


def hw_counted(workload):
    """Hardware-counted figures of the dominant kernel from the committed ncu capture of this build
    (profiles/r2_ncu_summary.json, written by tools/summarise_ncu.py); None when there is none."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_summary.json")))[workload]
        return rec
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--resamples", type=int, default=0, help="per GPU per step (c3: per step in total)")
    ap.add_argument("--mapping", default="auto", choices=["auto", "thread", "warp"])
    ap.add_argument("--schedule", default="auto", choices=["auto", "refill", "rounds"])
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
    import oracle_lib as oracle  # cpu_baseline legs and the c3 ensemble check only

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: fitter_mc_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    t_init0 = time.perf_counter()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    t_rendezvous = time.perf_counter() - t_init0

    wl = WORKLOADS[args.workload]
    strong = wl["scaling"] == "strong"
    model, x, y, err = make_problem(args.workload)
    ndata = len(x)
    info = fmc.initialise_model(model)
    p, q = info["nparam"], info["nprop"]
    t0 = time.perf_counter()
    ctx = fmc.Context(local_rank)
    ctx.set_data(model, x, y, err)
    t_context = time.perf_counter() - t0
    if args.mapping != "auto":
        ctx.set_fit_mapping(args.mapping)
    if args.schedule != "auto":
        ctx.set_schedule(args.schedule)
    # the library's own communicator: rank 0's id goes round by torch.distributed, then every
    # all-reduce of the run is fmc_allreduce (one ncclAllReduce on the kernels' stream)
    t0 = time.perf_counter()
    if world > 1:
        box = [fmc.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.nccl_init_rank(box[0], world, rank)
    t_nccl = time.perf_counter() - t0
    # every rank derives the same start from the same data on its own GPU: offset_data(.FALSE.); fit_data
    zeros = torch.zeros(ndata, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    ctx.launch(info["params_init"], 1, z_inject_dev=zeros.data_ptr(), keep=True)
    start = ctx.finish()["params"][0]
    if wl["dataset"] == "c5":
        import datasets

        start = datasets.C5_FITTED.copy()  # committed: the initial fit of the stress data is chaotic (tests/datasets.py)

    if strong:
        total = args.resamples or wl["total"]
        lo, hi = total * rank // world, total * (rank + 1) // world
        nres, step_ids = hi - lo, total
    else:
        nres = args.resamples or wl["resamples"]
        lo, step_ids = rank * nres, nres * world
    # a dedicated (non-default) torch stream carries the fit kernels, the all-reduce and the
    # timing events, so the CUDA events see exactly the work they bracket
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def device_step(step_no):
        """kernels + all-reduce with the data resident; returns the step's device time (ms)."""
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.launch(start, nres, seed=SEED, first_counter=step_no * step_ids + lo)
        ctx.allreduce()  # the single collective of the path: 31 doubles over NVLink
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

    # nvidia-smi takes a while to deliver its first sample: started before the warm-up, counted only
    # inside the timed region (20 ms period)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    first_ms = device_step(0)
    t_first_step = time.perf_counter() - t0
    for w in range(1, args.warmup):
        device_step(w)

    sync_all()
    sampler.begin()
    t_wall0 = time.perf_counter()
    ms = [device_step(args.warmup + k) for k in range(args.steps)]
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = float(sum(ms))
    kernel_ms = ctx.kernel_ms()
    launches_per_step = ctx.launch_count()
    last_step = args.warmup + args.steps - 1
    last = fmc.decode_accumulator(model, start, ctx.read_accumulator())
    assert last["nfit"] == step_ids, (last["nfit"], step_ids)
    n_anomalies = ctx.anomalies()[2]  # fits of the last step that hit a solver safety cap (normally 0)

    # ---- e2e: public host API with HOST buffers; per step the inputs go host->device and the
    # (all-reduced) accumulator block comes back device->host, all inside the timed region and
    # all on the one stream that also carries the kernels and the collective.
    def e2e_step(step_no):
        ctx.set_data(model, x, y, err)              # H2D of the step's inputs (x, y, err)
        ctx.launch(start, nres, seed=SEED, first_counter=step_no * step_ids + lo)  # H2D of params_start
        ctx.allreduce()
        host_block = ctx.read_accumulator()         # D2H of the result block (synchronises)
        ctx.wait()
        return host_block

    e2e_step(1000)
    sync_all()
    t0 = time.perf_counter()
    for k in range(args.steps):
        last_block = e2e_step(2000 + k)
    sync_all()
    e2e_s = time.perf_counter() - t0
    e2e_check = fmc.decode_accumulator(model, start, last_block)
    assert e2e_check["nfit"] == step_ids, (e2e_check["nfit"], step_ids)

    # max over ranks
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, t_wall, t_context, t_nccl, t_first_step], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, t_wall, t_context, t_nccl, t_first_step = [float(v) for v in t.cpu()]
    total_fits = float(step_ids) * args.steps
    value = total_fits / (dev_ms * 1e-3)
    e2e_value = total_fits / e2e_s

    checks = {}
    if strong and rank == 0:
        # (1) the N-GPU sums of the last timed step against ONE GPU running the same global ids alone
        if world > 1:
            ctx.launch(start, step_ids, seed=SEED, first_counter=last_step * step_ids)
            ctx.wait()
            alone = fmc.decode_accumulator(model, start, ctx.read_accumulator())
            rel = max(float(np.max(np.abs(alone[k] - last[k]) / np.abs(alone[k])))
                      for k in ("sum_p", "sum_p2", "sum_prop", "sum_prop2"))
            checks["n_gpu_vs_1_gpu_sums_max_rel"] = rel
            checks["n_gpu_vs_1_gpu_sums_ok"] = bool(rel <= 1e-12 and np.array_equal(alone["rc_count"], last["rc_count"]))
        # (2) north-star check 2 at the target size: ensemble means and errors of the 1e8-resample run against
        # the reference's own serial RANLUX+polar stream (oracle, nb resamples), 3 combined standard errors
        nb = int(os.environ.get("FMC_C3_ORACLE_SAMPLE", 8000))
        oracle.initialise_rng()
        ref = oracle.bootstrap(model, x, y, err, start, nb, per_resample=False, nthreads=os.cpu_count() or 1)
        rm, rs = oracle.finalise_moments(nb, ref["sum_p"], ref["sum_p2"])
        gm, gs = fmc.finalise_moments(step_ids, last["sum_p"], last["sum_p2"])
        zm = (gm - rm) / np.sqrt(gs ** 2 / step_ids + rs ** 2 / nb)
        zs = (gs - rs) / (rs * np.sqrt(0.5 / step_ids + 0.5 / nb))
        checks.update({"ensemble_vs_ranlux_oracle": {"n_gpu": step_ids, "n_oracle": nb, "mean": gm.tolist(),
                                                     "err": gs.tolist(), "oracle_mean": rm.tolist(), "oracle_err": rs.tolist(),
                                                     "mean_dev_in_se": zm.tolist(), "err_dev_in_se": zs.tolist()},
                       "ensemble_ok": bool((np.abs(zm) <= 3.0).all() and (np.abs(zs) <= 3.5).all())})
    if world > 1:
        sync_all()
        dist.destroy_process_group()

    if rank == 0:
        # FP64 roofline denominator: DFMA-chain microbenchmark on this GPU, after the timed work
        peak_tflops, _ = fmc.measure_dfma_peak(local_rank, 20000)
        tot = last["totals"].astype(float) / max(last["nfit"], 1)
        F, G = tot[0], tot[1]
        A = max(G - 2.0, 0.0)
        w_fit = algorithmic_flops_per_fit(ndata, p, F, G, A, wl["c_m"])
        achieved = w_fit * (nres / (kernel_ms * 1e-3)) / 1e12
        hw = hw_counted(args.workload)
        rc = last["rc_count"]
        line = {
            "metric": wl["metric"], "value": value, "unit": "fits/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["text"]},
            "details": {"resamples_per_gpu_per_step": nres, "resamples_per_step": step_ids, "seed": SEED,
                        "jacobian": "forward differences (nl2sno, the reference's path)",
                        "l2": "flushed between steps (256 MiB memset); the data rows live in shared memory",
                        "geometry": ctx.geometry(), "schedule": ctx.schedule(),
                        "per_fit": {"nfcall": F, "ngcall": G, "niter": tot[2], "passes": tot[3]},
                        "return_codes_3_to_10_last_step": rc[3:11].tolist(),
                        "non_converged_fraction_last_step": float(rc[7:11].sum()) / max(last["nfit"], 1),
                        "wall_s_timed_region": t_wall, "solver_safety_cap_hits_last_step": int(n_anomalies),
                        "collective": "fmc_allreduce: one ncclAllReduce of %d doubles per step inside the C ABI (NCCL %s)"
                                      % (ctx.accumulator_len(), fmc.nccl_available()),
                        "one_time_costs_s": {"torch_rendezvous": t_rendezvous, "context_and_h2d": t_context,
                                             "nccl_comm_init": t_nccl, "first_step_incl_module_load": t_first_step,
                                             "first_step_device_ms": first_ms}},
            "e2e": {"value": e2e_value, "unit": "fits/s", "h2d_bytes_per_step": int(8 * (3 * ndata + p)),
                    "d2h_bytes_per_step": int(8 * ctx.accumulator_len())},
            "gpu_launches": int(launches_per_step * args.steps * world),  # counted by the library (fmc_last_launch_count)
            "clocks": clocks,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / peak_tflops,
                         "hw_counted_tflops": hw.get("fp64_tflops") if hw else None,
                         "traffic": hw.get("dram_bytes_per_launch") if hw else None,
                         "ncu_source": hw.get("source") if hw else None,
                         "note": "FP64-FMA bound (SURVEY.md 8d), not hbm/tensor; peak = DFMA-chain microbenchmark "
                                 "measured in this run; achieved = %.0f algorithmic flops/fit x kernel fits/s; "
                                 "hw_counted_tflops / traffic = ncu counters of the dominant kernel, per launch, from "
                                 "the committed capture named in ncu_source; achieved counts the REFERENCE algorithm's "
                                 "operations (one Jacobian per gradient call it makes), so it exceeds the hardware count "
                                 "and can exceed the peak where the engine skips Jacobians nobody reads (round schedule "
                                 "with the residual sum first, DESIGN.md 2.8)" % w_fit},
        }
        if checks:
            line["checks"] = checks
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            legs = cpu_legs(args.workload, model, x, y, err, start, cores, ("reference", "lean", "solver_only"))
            line["cpu_baseline"] = legs["reference"]
            line["cpu_baseline_lean"] = legs["lean"]
            line["cpu_solver_only"] = legs["solver_only"]
            line["mcfit_py"] = mcfit_py_leg()
        RESULT_OUT.write(json.dumps(line) + "\n")
        RESULT_OUT.flush()
        if checks and not all(v for k, v in checks.items() if k.endswith("_ok")):
            raise SystemExit("bench.py: a correctness check of the %s run failed: %s" % (args.workload, checks))


if __name__ == "__main__":
    main()
