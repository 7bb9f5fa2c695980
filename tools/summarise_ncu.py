"""Turn an `ncu --set full` capture (gpurun_out/*.ncu-rep) into the figures bench.py reports beside the
algorithmic roofline: hardware-counted FP64 FLOP/s of the dominant kernel, its DRAM traffic per launch,
FP64 pipe utilisation, achieved occupancy and warp execution efficiency (BASELINE.json north star).

usage: python tools/summarise_ncu.py <workload> <capture.ncu-rep> [<text summary out>]
Updates profiles/r2_ncu_summary.json[<workload>] (read by bench.py) and writes the text summary."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed", "smsp__cycles_elapsed.avg.per_second",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
SCALE = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12,
         "msecond": 1e-3, "usecond": 1e-6, "second": 1.0, "nsecond": 1e-9, "Ghz": 1e9, "Mhz": 1e6, "hz": 1.0}


def num(d, u, k):
    return float(d[k].replace(",", "")) * SCALE.get(u.get(k, ""), 1.0)


def main():
    workload, rep = sys.argv[1], sys.argv[2]
    txt = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "r2_ncu_%s_summary.txt" % workload)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], dict(zip(rows[0], rows[1]))
    recs = [dict(zip(hdr, r)) for r in rows[2:]]
    lines = ["# ncu --set full --clock-control none capture: %s (workload %s)" % (os.path.basename(rep), workload),
             "# one block per profiled launch; FP64 FLOP/s = (2*dfma + dadd + dmul) thread instructions / duration"]
    best = None
    for d in recs:
        name = d["Kernel Name"].split("(")[0].replace("void ", "")
        lines += ["", "Kernel: " + d["Kernel Name"][:160]]
        for k in WANT:
            if k in d:
                lines.append("  %-84s %s %s" % (k, d[k], units.get(k, "")))
        try:
            per_cycle = (2 * float(d["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"].replace(",", ""))
                         + float(d["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"].replace(",", ""))
                         + float(d["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"].replace(",", "")))
            hz = num(d, units, "smsp__cycles_elapsed.avg.per_second")
            dur = num(d, units, "gpu__time_duration.sum")
            tflops = per_cycle * hz / 1e12
            dram = num(d, units, "dram__bytes_read.sum") + num(d, units, "dram__bytes_write.sum")
            lines.append("  derived: hardware-counted FP64 %.2f TFLOP/s (%.0f flop/cycle of 18944 nominal = %.1f%%, at %.3f GHz); "
                         "DRAM %.3e B per launch" % (tflops, per_cycle, 100 * per_cycle / 18944, hz / 1e9, dram))
            rec = {"kernel": name, "duration_ms": dur * 1e3, "fp64_tflops": tflops, "fp64_flop_per_cycle": per_cycle,
                   "sm_ghz": hz / 1e9, "dram_bytes_per_launch": dram,
                   "fp64_pipe_active_pct": float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]),
                   "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                   "achieved_occupancy_pct": float(d["sm__warps_active.avg.pct_of_peak_sustained_active"]),
                   "warp_execution_efficiency_threads_per_inst": float(d["smsp__thread_inst_executed_per_inst_executed.ratio"]),
                   "registers": int(d["launch__registers_per_thread"]),
                   "source": os.path.relpath(txt, ROOT)}
            if best is None or dur > best["duration_ms"] * 1e-3:
                best = rec
        except Exception as e:  # noqa: BLE001
            lines.append("  derived: n/a (%s)" % e)
    open(txt, "w").write("\n".join(lines) + "\n")
    path = os.path.join(ROOT, "profiles", "r2_ncu_summary.json")
    allrec = json.load(open(path)) if os.path.exists(path) else {}
    if best:
        try:
            best["build"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True,
                                           cwd=ROOT).stdout.strip()
        except Exception:
            pass
        allrec[workload] = best
        json.dump(allrec, open(path, "w"), indent=1)
    print("\n".join(lines[:60]))
    print(json.dumps(best, indent=1))


if __name__ == "__main__":
    main()
