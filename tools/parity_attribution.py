"""Which decision makes a fit of the product differ from the reference algorithm?  

Per-fit parity between the product's solver and the oracle's restatement of nl2sno is 1e-9 in the
median with a tail up to ~6e-8.  This tool attributes the tail.  For every fit of a sample it runs
the oracle and the HOST IMAGE of the product's solver (tests/host_emul: the same nl2_solver.cuh /
pass_eval.cuh / pert.cuh sources compiled by g++ with -DFMC_TRACE) on identical injected resamples,
records after EVERY call of `assess` (toms573.f90:1191) what the solver decided -- iv(irc), model
used, trial f, reldx, radius, pivot order of the QR/Cholesky factorisation -- and classifies the
FIRST record at which the two differ:

  same path            all decisions identical in both nl2sno calls: the deviation is pure arithmetic
  stop test            one side ends the call (irc >= 7: x-, relative-function-, absolute-function-,
                       singular or false convergence, toms573.f90:1905-1948) where the other takes a step
  accept / radius      both continue but assess returns a different irc (step accepted vs recomputed
                       with a smaller / larger radius, toms573.f90:1573-1904)
  keep / restore x     same irc, but one side keeps the trial point and the other goes back to the previous
                       one (iv(restor)): the sign of f(x) - f(x0) at rounding level (toms573.f90:1700-1760)
  model switch         same irc but the other quadratic model is used for the next step (1217-1235)
  pivot order          same irc and model, different column order out of qrfact (4436-4460) /
                       the pivoted Cholesky that replaces it
  extra record         one trace is a prefix of the other (one more trial evaluation on one side)

It does so for three builds of the host image: the product as shipped (perturbation arithmetic,
Cholesky of J^T J), -DFMC_FD_PLAIN (p + 1 plain evaluations and a subtraction, as nl2sno does), and
the analytic-Jacobian mode against the oracle's nl2sol + madj (no finite-difference noise at all).

usage: python tools/parity_attribution.py [config=c2] [nres=1024]  ->  JSON on stdout, table on stderr
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datasets  # noqa: E402
import oracle_lib as oracle  # noqa: E402

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
CAP, W = 96, 10
EMUL_DIR = os.path.join(ROOT, "tests", "host_emul")


def build_emul(tag, defines):
    so = os.path.join(EMUL_DIR, "libfmc_emul_%s.so" % tag)
    src = os.path.join(EMUL_DIR, "emul.cpp")
    csrc = os.path.join(ROOT, "fitter_mc_b200", "csrc")
    deps = [src] + [os.path.join(dp, f) for dp, _, fs in os.walk(csrc) for f in fs if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-x", "c++", "-std=c++17", "-fPIC", "-shared", "-w", "-DFMC_TRACE"]
                              + defines + ["-o", so, src])
    L = C.CDLL(so)
    L.emul_fit_trace.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _ip, _dp, C.c_int, _ip]
    return L


def d(a):
    return a.ctypes.data_as(_dp)


def oracle_trace(model, x, yfit, err, start, flags):
    L = oracle.lib()
    L.fmc_oracle_fit_trace.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_int, _ip, _dp, C.c_int, _ip]
    p = np.array(start, dtype=np.float64)
    cnt = np.zeros((2, 6), dtype=np.int32)
    tr = np.zeros((2, CAP, W))
    ln = np.zeros(2, dtype=np.int32)
    rc = L.fmc_oracle_fit_trace(model, len(x), d(x), d(yfit), d(err), d(p), flags, cnt.ctypes.data_as(_ip), d(tr), CAP,
                                ln.ctypes.data_as(_ip))
    assert rc == 0
    return p, cnt, tr, ln


def emul_trace(L, model, jac, x, yfit, err, start):
    p = np.array(start, dtype=np.float64)
    cnt = np.zeros((2, 5), dtype=np.int32)
    tr = np.zeros((2, CAP, W))
    ln = np.zeros(2, dtype=np.int32)
    rc = L.emul_fit_trace(model, jac, len(x), d(x), d(yfit), d(err), d(p), cnt.ctypes.data_as(_ip), d(tr), CAP,
                          ln.ctypes.data_as(_ip))
    assert rc == 0
    return p, cnt, tr, ln


def classify(tro, lno, tre, lne):
    """(category, call, index of the first differing record) for one fit."""
    for call in range(2):
        a, b = tro[call, :lno[call]], tre[call, :lne[call]]
        for i in range(min(len(a), len(b))):
            irc_a, irc_b = int(a[i, 0]), int(b[i, 0])
            if irc_a != irc_b:
                if (irc_a >= 7) != (irc_b >= 7):
                    return "stop test", call, i
                return "accept / radius", call, i
            if int(a[i, 8]) != int(b[i, 8]):
                return "keep / restore x", call, i
            if int(a[i, 1]) != int(b[i, 1]):
                return "model switch", call, i
            if int(a[i, 7]) != int(b[i, 7]) and int(a[i, 1]) % 2 == 1:   # pivots matter for the Gauss-Newton model
                return "pivot order", call, i
        if len(a) != len(b):
            return "extra record", call, min(len(a), len(b))
    return "same path", -1, -1


def run(name, nres, seed=2024):
    model, x, y, err = datasets.dataset(name)
    info = oracle.model_info(model)
    start, _ = oracle.fit_data(model, x, y, err, info["params_init"])
    z = np.random.default_rng(seed).normal(size=(nres, len(x)))
    variants = [("shipped: perturbation arithmetic + Cholesky of JtJ", build_emul("trace", []), 0, oracle.LSVMIN_PER_FIT),
                ("FMC_FD_PLAIN: p+1 plain evaluations + Cholesky of JtJ", build_emul("trace_plain", ["-DFMC_FD_PLAIN"]), 0,
                 oracle.LSVMIN_PER_FIT),
                ("analytic Jacobian both sides (no FD noise)", build_emul("trace", []), 1,
                 oracle.LSVMIN_PER_FIT | oracle.ANALYTIC_JAC)]
    out = {"config": name, "nres": nres, "variants": {}}
    for label, L, jac, flags in variants:
        cats, rels, where, last_rec = [], [], [], []
        for r in range(nres):
            yfit = y + z[r] * err
            po, co, tro, lno = oracle_trace(model, x, yfit, err, start, flags)
            pe, ce, tre, lne = emul_trace(L, model, jac, x, yfit, err, start)
            cat, call, idx = classify(tro, lno, tre, lne)
            cats.append(cat)
            rels.append(float(np.max(np.abs(pe - po) / np.abs(po))))
            where.append(call)
            # is the first difference at the LAST record of the call on either side (the noise-level last step)?
            last_rec.append(bool(call >= 0 and (idx >= lno[call] - 1 or idx >= lne[call] - 1)))
        rels, cats, where, last_rec = np.array(rels), np.array(cats), np.array(where), np.array(last_rec)
        rec = {"median": float(np.median(rels)), "p99": float(np.quantile(rels, 0.99)), "max": float(rels.max()),
               "frac_le_1e-9": float((rels <= 1e-9).mean()), "categories": {}}
        for cat in ("same path", "stop test", "accept / radius", "keep / restore x", "model switch", "pivot order",
                    "extra record"):
            sel = cats == cat
            if sel.any():
                rec["categories"][cat] = {"fits": int(sel.sum()), "frac": float(sel.mean()),
                                          "median": float(np.median(rels[sel])), "max": float(rels[sel].max()),
                                          "in_second_call": float((where[sel] == 1).mean()) if cat != "same path" else None,
                                          "at_last_record_of_call": float(last_rec[sel].mean()) if cat != "same path" else None}
        tail = rels > 1e-8
        rec["tail_gt_1e-8"] = {"fits": int(tail.sum()),
                               "by_category": {c: int((cats[tail] == c).sum()) for c in set(cats[tail])}}
        out["variants"][label] = rec
        print("\n[%s] %s: median %.2e p99 %.2e max %.2e frac<=1e-9 %.3f" %
              (name, label, rec["median"], rec["p99"], rec["max"], rec["frac_le_1e-9"]), file=sys.stderr)
        for cat, v in rec["categories"].items():
            print("   %-16s %5d fits (%.3f)  median %.2e  max %.2e  2nd call %s  last record %s" %
                  (cat, v["fits"], v["frac"], v["median"], v["max"], v["in_second_call"], v["at_last_record_of_call"]),
                  file=sys.stderr)
        print("   tail (> 1e-8): %s" % rec["tail_gt_1e-8"], file=sys.stderr)
    return out


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    nres = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    print(json.dumps(run(name, nres), indent=1))
