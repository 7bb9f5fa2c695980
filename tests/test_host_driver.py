"""Row f1: the host driver pieces restated from the reference -- reader,
sort, report formatting -- and, on a GPU, the whole `./fitter_mc datafile.dat` transcript."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest
from scipy.special import gammaincc

import datasets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fitter_mc_b200")


@pytest.fixture(scope="module")
def host():
    subprocess.check_call(["make", "-C", PKG, "libfmc_host.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(PKG, "libfmc_host.so"))
    L.fmc_host_gammq.restype = C.c_double
    L.fmc_host_gammq.argtypes = [C.c_double, C.c_double]
    L.fmc_host_write_mean.argtypes = [C.c_double, C.c_double, C.c_int, C.c_char_p]
    L.fmc_host_list_directed_real.argtypes = [C.c_double, C.c_char_p]
    L.fmc_host_isdataline.argtypes = [C.c_char_p]
    L.fmc_host_wordwrap.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    dp = C.POINTER(C.c_double)
    L.fmc_host_read_data.argtypes = [C.c_char_p, dp, dp, dp, C.c_int, C.POINTER(C.c_int), C.c_char_p]
    return L


def _wm(host, av, err, prec=1):
    buf = C.create_string_buffer(72)
    host.fmc_host_write_mean(av, err, prec, buf)
    return buf.value.decode()


def test_write_mean_notation(host):
    """utils.f90:128-218, e.g. 0.123546(7)."""
    assert _wm(host, 0.123546, 0.000007) == "0.123546(7)"
    assert _wm(host, 0.5039360, 0.0126) == "0.50(1)"
    assert _wm(host, 1.96044, 0.0351) == "1.96(4)"
    assert _wm(host, -2.5, 0.3) == "-2.5(3)"
    assert _wm(host, 1234.5, 20.0) == "1230(20)"
    assert _wm(host, 0.00012, 0.00096) == "0.000(1)"      # error rounds up to the next figure
    assert _wm(host, 3.14159, 0.0123, 2) == "3.142(12)"
    assert _wm(host, 1.0, 0.0).startswith("ERROR: NON-POSITIVE ERROR BAR")


def test_gammq_matches_scipy(host):
    """utils.f90:239-335 against scipy.special.gammaincc (what mcfit.py uses, mcfit.py:73)."""
    for a, x in [(18.5, 24.08), (498.0, 520.0), (0.5, 0.1), (3.0, 10.0), (18.5, 5.0), (100.0, 99.0)]:
        assert np.isclose(host.fmc_host_gammq(a, x), gammaincc(a, x), rtol=1e-10), (a, x)
    # the probe file of BASELINE.md: reduced chi^2 1.3016073130752277 with 37 dof -> Q = 0.10353124942978763
    assert np.isclose(host.fmc_host_gammq(18.5, 0.5 * 37 * 1.3016073130752277), 0.10353124942978763, rtol=1e-9)


def test_list_directed_real_layout(host):
    buf = C.create_string_buffer(32)
    for v in (1.3016073130752208, 0.5039360439408515, -2.4643745968509885, 123456.789, 1e-5, 3e20, 0.0):
        host.fmc_host_list_directed_real(v, buf)
        s = buf.value.decode()
        assert len(s) == 25 and float(s) == pytest.approx(v, rel=1e-15)
    host.fmc_host_list_directed_real(1.3016073130752208, buf)
    assert buf.value.decode() == "  1.3016073130752208     "   # G25.17E3: F20.16 + 5 blanks
    host.fmc_host_list_directed_real(1e-5, buf)
    assert buf.value.decode().strip() == "1.0000000000000001E-005"


def test_isdataline_and_wordwrap(host):
    for s, want in [(b"1.0 2.0 0.1", 1), (b"  -3.5d0 1 1", 1), (b"# comment", 0), (b"x y err", 0), (b"", 0),
                    (b".5,1,1", 1), (b"1e3 2 3", 1), (b"abc 1 2", 0)]:
        assert host.fmc_host_isdataline(s) == want, s
    out = C.create_string_buffer(4096)
    text = b"Reading data from " + b"a_rather_long_directory_name/" * 4 + b"data file.dat."
    host.fmc_host_wordwrap(text, 79, out, 4096)
    lines = out.value.decode().split("\n")[:-1]
    assert all(len(l) <= 80 and l.startswith(" ") for l in lines) and len(lines) >= 2
    assert "".join(l[1:] for l in lines).replace(" ", "") == text.decode().replace(" ", "")


def test_reader_sorts_and_rejects(host, tmp_path):
    dp = C.POINTER(C.c_double)
    x, y, err = datasets.c1_probe()
    p = tmp_path / "probe.dat"
    order = np.random.default_rng(0).permutation(len(x))
    with open(p, "w") as f:
        f.write("# x y err\nheader line without numbers\n")
        for i in order:
            f.write("%.18e  %.18e, %.18e\n" % (x[i], y[i], err[i]))
        f.write("\n")
    X, Y, E = np.zeros(100), np.zeros(100), np.zeros(100)
    flags, msg = C.c_int(0), C.create_string_buffer(256)
    n = host.fmc_host_read_data(str(p).encode(), X.ctypes.data_as(dp), Y.ctypes.data_as(dp), E.ctypes.data_as(dp),
                                100, C.byref(flags), msg)
    assert n == 40 and flags.value & 1                      # "raw x data are not in ascending order"
    assert np.array_equal(X[:40], x) and np.array_equal(Y[:40], y) and np.array_equal(E[:40], err)
    bad = tmp_path / "bad.dat"
    bad.write_text("1 2 0.1\n2 3 0.0\n")
    assert host.fmc_host_read_data(str(bad).encode(), X.ctypes.data_as(dp), Y.ctypes.data_as(dp),
                                   E.ctypes.data_as(dp), 100, C.byref(flags), msg) == -1
    assert b"non-positive error bar" in msg.value            # bootstrap.f90:134-138
    assert host.fmc_host_read_data(b"/nonexistent", X.ctypes.data_as(dp), Y.ctypes.data_as(dp), E.ctypes.data_as(dp),
                                   100, C.byref(flags), msg) == -1


@pytest.mark.gpu
def test_fitter_mc_transcript_on_the_probe_file(tmp_path, oracle):
    """`./fitter_mc datafile.dat`: the reference's transcript with the
    numbers of BASELINE.md section 2."""
    subprocess.check_call(["make", "-C", PKG, "fitter_mc"], stdout=subprocess.DEVNULL)
    x, y, err = datasets.c1_probe()
    np.savetxt(tmp_path / "probe.dat", np.c_[x, y, err])
    out = subprocess.run([os.path.join(PKG, "fitter_mc"), "probe.dat"], cwd=tmp_path, capture_output=True, text=True,
                         timeout=120)
    assert out.returncode == 0, out.stderr
    t = out.stdout
    for needle in (" FITTER_MC\n =========", " Reading data from probe.dat.", " Number of data lines: 40",
                   " Model used is: test", " Number of parameters: 3", " Initial parameter set:",
                   " Fitted parameter set:", " Monte Carlo fitted parameter set (using 100000 random points):",
                   "   Histogram of ab written to histogram_ab.dat.", " Program finished."):
        assert needle in t, needle
    nums = [float(v) for v in re.findall(r"Parameter \S+\s+=\s+([-0-9.E+]+)", t)]
    fitted = nums[3:6]
    assert np.allclose(fitted, [0.5039360403603528, 1.960438556490636, 1.2822067222414264], rtol=5e-8)
    chi = [float(v) for v in re.findall(r"Reduced chi\^2 function =\s+([-0-9.E+]+)", t)]
    assert np.isclose(chi[1], 1.3016073130752277, rtol=1e-10)
    q = [float(v) for v in re.findall(r"this bad =\s+([-0-9.E+]+)", t)]
    assert np.isclose(q[1], 0.10353124942978763, rtol=1e-8)
    mc = re.findall(r"Parameter (\S+)\s+=\s+([-0-9.E+]+)\s+\+/-\s+([-0-9.E+]+)", t)
    assert [m[0] for m in mc] == ["a", "b", "al"]
    want_m, want_s = [0.50388, 1.96112, 1.28296], [0.01247, 0.03474, 0.04703]   # mcfit.py, 2000 samples
    for (_, mean, sd), wm, ws in zip(mc, want_m, want_s):
        assert abs(float(mean) - wm) < 4 * ws / np.sqrt(2000) and abs(float(sd) - ws) < 4 * ws / np.sqrt(4000)
    assert re.search(r"=\s+0\.50\(1\)", t) and re.search(r"Property  ab =", t)
    h = np.loadtxt(tmp_path / "histogram_ab.dat")
    assert h.shape == (100000,) and abs(h.mean() - 2.4650) < 0.001


@pytest.mark.gpu
def test_fitter_mc_binned_histogram_mode(tmp_path):
    """FMC_HISTOGRAM_BINS: histogram_<prop>.dat holds bin centres and counts from the device-side
    binning (row f4) instead of one line per resample; same transcript, same moments."""
    subprocess.check_call(["make", "-C", PKG, "fitter_mc"], stdout=subprocess.DEVNULL)
    x, y, err = datasets.c1_probe()
    np.savetxt(tmp_path / "probe.dat", np.c_[x, y, err])
    env = dict(os.environ, FMC_HISTOGRAM_BINS="400", FMC_NRAND="200000")
    out = subprocess.run([os.path.join(PKG, "fitter_mc"), "probe.dat"], cwd=tmp_path, capture_output=True, text=True,
                         timeout=120, env=env)
    assert out.returncode == 0, out.stderr
    t = out.stdout
    assert " Monte Carlo fitted parameter set (using 200000 random points):" in t
    assert "   Histogram of ab written to histogram_ab.dat." in t
    head = open(tmp_path / "histogram_ab.dat").readline()
    assert head.startswith("# binned on the GPU: 400 bins on [")
    below, above = (int(v) for v in re.search(r"below: (\d+); above: (\d+)", head).groups())
    h = np.loadtxt(tmp_path / "histogram_ab.dat")
    assert h.shape == (400, 2) and int(h[:, 1].sum()) + below + above == 200000
    assert below + above < 20                                   # +/- 8 sigma
    mean = (h[:, 0] * h[:, 1]).sum() / h[:, 1].sum()
    mc = re.findall(r"Property  ab =\s+([-0-9.E+]+)\s+\+/-\s+([-0-9.E+]+)", t)
    assert abs(mean - float(mc[0][0])) < 2e-4                    # binned mean vs the exact moment sums
    qs = [float(v) for v in re.search(r"Quantiles of ab.*\n\s+(.*)\n", t).group(1).split()]
    assert len(qs) == 5 and np.all(np.diff(qs) > 0)
    sd = float(mc[0][1])
    assert abs((qs[3] - qs[1]) / 2 - sd) < 0.05 * sd and abs(qs[2] - float(mc[0][0])) < 0.05 * sd


@pytest.mark.gpu
def test_fitter_mc_analytic_jacobian_switch(tmp_path, oracle):
    """FMC_JACOBIAN=analytic: the README's "replace the call to nl2sno with a call to nl2sol"."""
    subprocess.check_call(["make", "-C", PKG, "fitter_mc"], stdout=subprocess.DEVNULL)
    x, y, err = datasets.c1_probe()
    np.savetxt(tmp_path / "probe.dat", np.c_[x, y, err])
    env = dict(os.environ, FMC_JACOBIAN="analytic", FMC_NRAND="20000", FMC_HISTOGRAM="0")
    out = subprocess.run([os.path.join(PKG, "fitter_mc"), "probe.dat"], cwd=tmp_path, capture_output=True, text=True,
                         timeout=120, env=env)
    assert out.returncode == 0, out.stderr
    nums = [float(v) for v in re.findall(r"Parameter \S+\s+=\s+([-0-9.E+]+)", out.stdout)]
    want, _ = oracle.fit_data(0, x, y, err, [0.0, 2.0, 1.0], flags=oracle.LSVMIN_PER_FIT | oracle.ANALYTIC_JAC)
    assert np.allclose(nums[3:6], want, rtol=1e-8)
    assert "Histogram of" not in out.stdout and " Program finished." in out.stdout
