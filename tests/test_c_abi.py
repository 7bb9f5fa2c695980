"""CPU-side checks of the drop-in boundary: libfmc_b200.so loads, exports every symbol that
include/fmc_b200.h declares, serves the model plug-in's host images, decodes accumulator
blocks, and fails loudly (no CPU fallback) when asked to compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import datasets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fmc():
    import fitter_mc_b200

    fitter_mc_b200.lib()
    return fitter_mc_b200


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "fmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fmc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(fmc):
    lib = ctypes.CDLL(fmc.lib_path())
    names = _declared_symbols()
    assert len(names) >= 20 and "fmc_bootstrap_run" in names and "fmc_fit_once" in names
    for n in names:
        assert hasattr(lib, n), "include/fmc_b200.h declares %s but the library does not export it" % n


def test_fortran_binding_names_match_the_header():
    f90 = open(os.path.join(ROOT, "fitter_mc_b200", "fortran", "fmc_b200_bindings.f90")).read()
    bound = set(re.findall(r"name='(fmc_[a-z0-9_]+)'", f90))
    assert bound and bound <= set(_declared_symbols())


def test_model_plugin_host_images_match_the_oracle(fmc, oracle):
    assert fmc.model_count() == oracle.lib().fmc_oracle_model_count()
    for model in range(fmc.model_count()):
        a, b = fmc.initialise_model(model), oracle.model_info(model)
        assert (a["name"], a["nparam"], a["nprop"], a["param_name"], a["prop_name"]) == \
               (b["name"], b["nparam"], b["nprop"], b["param_name"], b["prop_name"])
        assert np.array_equal(a["params_init"], b["params_init"])
        p = datasets.TRUTH[model]
        xs = np.linspace(0.2, 9.0, 7)
        assert np.allclose(fmc.model_y(model, p, xs), oracle.model_y(model, p, xs), rtol=1e-15)
        assert np.allclose(fmc.model_y(model, p, xs), datasets.model_y(model, p, xs), rtol=1e-14)
        assert np.allclose(fmc.derived_properties(model, p), oracle.derived_properties(model, p), rtol=1e-15)
    # the reference's default model (bootstrap.f90:373-386, 395, 406)
    d = fmc.initialise_model(0)
    assert d["name"] == "test" and d["param_name"] == ["a ", "b ", "al"] and d["prop_name"] == ["ab"]
    assert list(d["params_init"]) == [0.0, 2.0, 1.0]


def test_decode_accumulator_recovers_raw_sums(fmc):
    from fitter_mc_b200 import distributed as fd

    rng = np.random.default_rng(3)
    start = np.array([0.5, 2.0, 1.3])
    params = start + 1e-3 * rng.normal(size=(5000, 3))       # sigma/|mean| ~ 1e-3: cancellation-prone
    props = (params[:, 0] + params[:, 1])[:, None]
    acc = fd.encode_accumulator(params, props, start, fmc.derived_properties(0, start), rc=np.full(5000, 4))
    assert len(acc) == fd.accumulator_layout(3, 1)["length"]
    dec = fmc.decode_accumulator(0, start, acc)
    assert dec["nfit"] == 5000 and dec["rc_count"][4] == 5000
    assert np.allclose(dec["sum_p"], params.sum(0), rtol=1e-14)
    assert np.allclose(dec["sum_p2"], (params ** 2).sum(0), rtol=1e-14)
    mean, sd = fmc.finalise_moments(5000, dec["sum_p"], dec["sum_p2"])
    assert np.allclose(mean, params.mean(0), rtol=1e-14) and np.allclose(sd, params.std(0, ddof=1), rtol=1e-6)


def test_moments_formula_is_the_references(fmc, oracle):
    rng = np.random.default_rng(4)
    v = rng.normal(3.0, 0.1, size=(1000, 2))
    a = fmc.finalise_moments(1000, v.sum(0), (v * v).sum(0))
    b = oracle.finalise_moments(1000, v.sum(0), (v * v).sum(0))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    one = fmc.finalise_moments(1, v[0], v[0] ** 2)          # nrand_points = 1 branch, bootstrap.f90:338-340
    assert (one[1] == 0).all()


def test_no_cpu_fallback(fmc):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    model, x, y, err = datasets.dataset("c1")
    with pytest.raises(fmc.FmcError) as e:
        fmc.fit_data(x, y, err, [0.0, 2.0, 1.0], model)
    assert e.value.code == -1  # FMC_ERR_NO_DEVICE
    with pytest.raises(fmc.FmcError):
        fmc.monte_carlo_fit_data(x, y, err, model, nrand_points=8, initial_fit=False)
    with pytest.raises(fmc.FmcError):
        fmc.generate_offsets(1, 0, 4, 8)
    with pytest.raises(fmc.FmcError):
        fmc.Context(0)


def test_product_does_not_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under fitter_mc_b200/ or include/ may use it."""
    for base in ("fitter_mc_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".f90", ".cpp", "Makefile")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "oracle_lib" not in text and "fmc_oracle" not in text and "oracle/" not in text, \
                        os.path.join(dirpath, f)


def test_header_is_plain_c_and_a_c_program_links(fmc, tmp_path):
    """The boundary is a C ABI: include/fmc_b200.h must compile as C99 (no C++, no torch types) and a
    C program must link against libfmc_b200.so and call the host-side entry points."""
    import subprocess

    src = tmp_path / "use_abi.c"
    src.write_text('#include "fmc_b200.h"\n#include <stdio.h>\n'
                   'int main(void) {\n'
                   '  int np = 0, nq = 0; double init[FMC_MAX_PARAM]; char pn[FMC_MAX_PARAM][3], qn[FMC_MAX_PARAM][3], name[32];\n'
                   '  if (fmc_initialise_model(0, &np, &nq, init, pn, qn, name) != FMC_OK) return 1;\n'
                   '  printf("%s %d %d %.1f %s\\n", name, np, nq, fmc_model_y(0, init, 0.0), fmc_strerror(FMC_ERR_NO_DEVICE));\n'
                   '  return fmc_model_count() < 1;\n}\n')
    exe = tmp_path / "use_abi"
    libdir = os.path.dirname(fmc.lib_path())
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror",
                           "-I", os.path.join(ROOT, "include"), str(src), "-L", libdir, "-lfmc_b200",
                           "-Wl,-rpath," + libdir, "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0
    assert out.stdout.startswith("test 3 1 2.0 no usable CUDA device")
