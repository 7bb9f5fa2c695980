"""Attribute the warp-stall samples and executed instructions of one kernel in an `ncu --set full
--import-source on` capture to SOURCE FUNCTIONS (the solver's routines are all inlined into the kernel, so
ncu's own per-function view shows one entry).

usage: python tools/ncu_region_breakdown.py <capture.ncu-rep> <kernel regex> <object.o | lib.so> [<launch id>]

Joins `ncu --page source --print-source sass --csv` (per SASS instruction: samples, executed instructions,
thread instructions, stall reasons) with `nvdisasm -g` of the same build (per SASS instruction: file and
line), then maps lines of nl2_solver.cuh to the routine they are in."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def function_ranges(path):
    """[(first line, name)] of the routines in a header, by their definition lines."""
    out = []
    pat = re.compile(r"^(?:template.*>\s*)?(?:FMC_HD|FMC_INLINE|__device__|inline|static)[^;(]*?\b([A-Za-z_0-9:<>]+)\s*\(")
    with open(path) as f:
        for no, line in enumerate(f, 1):
            if line.startswith((" ", "\t", "//", "#")):
                continue
            m = pat.match(line)
            if m and not line.rstrip().endswith(";"):
                out.append((no, m.group(1).split("::")[-1]))
    return out


def main():
    rep, kregex, obj = sys.argv[1], sys.argv[2], sys.argv[3]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k",
                          "regex:" + kregex], capture_output=True, text=True).stdout
    # several kernels / launches may follow each other: take the first (or the one asked for)
    blocks, cur = [], None
    for line in src.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = [line]
            blocks.append(cur)
        elif cur is not None:
            cur.append(line)
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    blk = blocks[which]
    kname = next(csv.reader([blk[0]]))[1]
    rows = list(csv.reader(io.StringIO("\n".join(blk[1:]))))
    hdr, rows = rows[0], rows[1:]
    col = {n: i for i, n in enumerate(hdr)}
    base = int(rows[0][col["Address"]], 16)

    # line info of the same build
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
        line_of = {}
        short = re.search(r"(\w+)\s*[<(]", kname.replace("void ", "")).group(1)
        for cub in sorted(os.listdir(td)):
            dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
            sec, want, f_l = None, False, None
            cands = collections.OrderedDict()
            for line in dis.splitlines():
                if line.startswith("\t.section\t.text."):
                    sec = line.split(".text.")[1].split(",")[0]
                    want = short in sec
                    if want:
                        cands[sec] = {}
                    continue
                if not want:
                    continue
                m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', line)
                if m:
                    f_l = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", line)
                if m:
                    cands[sec][int(m.group(1), 16)] = f_l
            # the instantiation whose instruction count matches the capture
            for sec, d in cands.items():
                if len(d) == len(rows):
                    line_of = d
        if not line_of:
            sys.exit("no disassembled function with %d instructions matches %s" % (len(rows), kname))

    ranges = {}
    for h in ("nl2_solver.cuh", "round_kernels.cuh", "bootstrap_kernel.cuh", "pass_eval.cuh", "pert.cuh", "rng.cuh",
              "fast_math.cuh"):
        p = os.path.join(ROOT, "fitter_mc_b200", "csrc", h)
        if os.path.exists(p):
            ranges[h] = function_ranges(p)

    def region(fl):
        if fl is None:
            return "?"
        f, l = fl
        name = ""
        for first, n in ranges.get(f, []):
            if first <= l:
                name = n
        return f.replace(".cuh", "") + ":" + name if name else f

    agg = collections.defaultdict(lambda: collections.Counter())
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    for r in rows:
        off = int(r[col["Address"]], 16) - base
        reg = region(line_of.get(off))
        a = agg[reg]
        a["sass"] += 1
        a["samples"] += int(r[col["# Samples"]] or 0)
        a["inst"] += int(r[col["Instructions Executed"]] or 0)
        a["thread_inst"] += int(r[col["Thread Instructions Executed"]] or 0)
        for s in stall_cols:
            a[s] += int(r[col[s]] or 0)
    tot_s = sum(a["samples"] for a in agg.values()) or 1
    tot_i = sum(a["inst"] for a in agg.values()) or 1
    print("# %s\n# %d SASS instructions, %d samples, %d warp instructions executed" % (kname, len(rows), tot_s, tot_i))
    print("# %-34s %6s %8s %8s %9s  top stall reasons" % ("region (file:routine)", "SASS", "samples", "inst", "lanes/inst"))
    for reg, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
        if a["samples"] * 500 < tot_s and a["inst"] * 500 < tot_i:
            continue
        st = sorted(((a[s], s[6:]) for s in stall_cols), reverse=True)[:3]
        print("  %-34s %6d %7.1f%% %7.1f%% %9.1f  %s" % (reg, a["sass"], 100.0 * a["samples"] / tot_s, 100.0 * a["inst"] / tot_i,
              a["thread_inst"] / max(a["inst"], 1), ", ".join("%s %.0f%%" % (n, 100.0 * v / max(a["samples"], 1)) for v, n in st)))


if __name__ == "__main__":
    main()
