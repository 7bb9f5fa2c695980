"""Turn the ncu artefacts of a round (gpurun_out/) into the text summaries kept under profiles/."""
import collections, csv, json, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = []
rows = list(csv.reader(open("gpurun_out/launches_%s.csv" % tag)))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[h], rows[h + 1:]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in data:
    agg.setdefault(r[ik].split("(")[0].replace("void ", ""), []).append(float(r[iv].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
out += ["# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (round %s)" % tag,
        "# ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are serialised and",
        "# cold-cache, so compare SHARES with the bench line, not absolutes",
        "# kernel | launches | total ms | share | mean ms"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    out.append("%s | %d | %.3f | %.1f%% | %.3f" % (k, len(v), sum(v) / 1e6, 100 * sum(v) / tot, sum(v) / len(v) / 1e6))
open("profiles/%s_bench_launches_summary.txt" % tag, "w").write("\n".join(out) + "\n")
print("\n".join(out))

# full capture: one block per profiled launch
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_%s_pipeline.ncu-rep" % tag, "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed", "smsp__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
lines = ["# ncu --set full --clock-control none capture of the pipeline kernels inside bench.py (round %s)" % tag,
         "# one block per profiled launch; FP64 FLOP/s = (2*dfma + dadd + dmul) / duration"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    lines.append("")
    for k in want:
        if k in d:
            lines.append("%-82s %s %s" % (k, d[k], u.get(k, "")))
    try:
        per_cycle = 2 * float(d["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"]) + float(
            d["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"]) + float(
            d["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"])
        hz = float(d["smsp__cycles_elapsed.avg.per_second"]) * {"Ghz": 1e9, "Mhz": 1e6, "hz": 1.0}.get(
            u["smsp__cycles_elapsed.avg.per_second"], 1e9)
        lines.append("%-82s %.2f TFLOP/s (%.0f flop/cycle of 18944 peak = %.1f%%, at %.3f GHz)" % (
            "derived: hardware-counted FP64 throughput", per_cycle * hz / 1e12, per_cycle, 100 * per_cycle / 18944, hz / 1e9))
    except Exception as e:
        lines.append("derived: n/a (%s)" % e)
open("profiles/%s_pipeline_ncu_summary.txt" % tag, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:70]))
