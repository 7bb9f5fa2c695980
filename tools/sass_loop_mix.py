"""Static instruction mix of the hot (row) loop of a compiled kernel: finds the backward branches in
the SASS of one function and prints the opcode histogram of the largest loop bodies.
usage: python tools/sass_loop_mix.py <object-or-so> <substring of the mangled kernel name> [min_instr]"""
import collections
import re
import subprocess
import sys


def sass(obj, sub):
    names = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True).stdout
    funs = [m.group(1) for m in re.finditer(r"Function (\S+):", names) if sub in m.group(1)]
    if not funs:
        raise SystemExit("no function matches " + sub)
    out = {}
    for f in funs:
        out[f] = subprocess.run(["cuobjdump", "-sass", "-fun", f, obj], capture_output=True, text=True).stdout
    return out


def loops(text):
    ins = [(int(m.group(1), 16), m.group(2).strip())
           for m in re.finditer(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", text, re.M)]
    res = []
    for a, t in ins:
        m = re.search(r"\bBRA\S*\s+(?:!?U?P\d+,\s*)*(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            res.append((int(m.group(1), 16), a))
    return ins, res


def mix(ins, lo, hi):
    c = collections.Counter()
    for a, t in ins:
        if lo <= a <= hi:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            op = t.split()[0]
            op = op if op.startswith("MUFU") else op.split(".")[0]
            c[op] += 1
    return c


if __name__ == "__main__":
    obj, sub = sys.argv[1], sys.argv[2]
    minn = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    for f, text in sass(obj, sub).items():
        ins, lp = loops(text)
        print(f, "instructions:", len(ins))
        # innermost loops only (no other backward branch inside), largest first
        inner = [(lo, hi) for lo, hi in lp if not any(lo <= l2 and h2 < hi for l2, h2 in lp if (l2, h2) != (lo, hi))]
        for lo, hi in sorted(set(inner), key=lambda ab: ab[0] - ab[1])[:6]:
            c = mix(ins, lo, hi)
            n = sum(c.values())
            if n < minn:
                continue
            fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
            print("  loop 0x%x..0x%x: %d instr, FP64 %d, LDL %d STL %d, RCP64H %d" %
                  (lo, hi, n, fp64, c["LDL"], c["STL"], c["MUFU.RCP64H"]))
            print("    ", dict(c.most_common(24)))
