"""Static guard on the compiled fit kernels (no GPU needed: cuobjdump reads the cubin inside
libfmc_b200.so): the shipped launch geometry is 256 threads x 1 block per SM, which only works while
the refill kernel stays within 255 registers WITHOUT spilling -- its stack frame must hold nothing but
the solver's own arrays.  A change that pushes the headline kernels over that edge fails here, before
anyone spends GPU time on it."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fitter_mc_b200", "libfmc_b200.so")


@pytest.fixture(scope="module")
def usage():
    if not shutil.which("cuobjdump") or not os.path.exists(LIB):
        pytest.skip("needs cuobjdump and the built library")
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, timeout=300).stdout
    res = {}
    for name, regs, stack in re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        res[name] = (int(regs), int(stack))
    assert res, out[:500]
    return res


def _find(usage, *parts):
    hits = [(k, v) for k, v in usage.items() if all(p in k for p in parts)]
    assert len(hits) == 1, (parts, [k for k, _ in hits])
    return hits[0][1]


def test_sm_100a_only(usage):
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True, timeout=300).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs                                  # written for B200 only, no other targets


# (the stress model is built with the compact, all-rolled solver -- csrc/kernels_model3.cu: 32 bytes more frame for
# the out-of-line helpers' arguments, still no spill in the row loop)
@pytest.mark.parametrize("model,solver_stack", [("9ModelUser", 1240), ("10ModelTest4", 1648), ("11ModelStiff4", 1680)])
def test_refill_kernels_of_small_models_do_not_spill(usage, model, solver_stack):
    # iterate_kernel<M, INJECT=0, SMEM=1, JAC_FD>: the kernel bench.py times
    regs, stack = _find(usage, "iterate_kernelINS_" + model + "ELb0ELb1ELi0EEEv")
    assert regs * 256 <= 65536
    assert stack == solver_stack                                     # the solver's arrays, no spill slots
    regs_a, stack_a = _find(usage, "iterate_kernelINS_" + model + "ELb0ELb1ELi1EEEv")   # analytic mode
    assert regs_a * 256 <= 65536 and stack_a == solver_stack


def test_uniform_kernels_are_light(usage):
    regs, stack = _find(usage, "init_shared_kernelINS_10ModelTest4ELb0E")
    assert regs <= 64 and stack <= 16
    regs, stack = _find(usage, "init_pass_kernelINS_10ModelTest4ELb0ELb1ELi0E")
    assert regs <= 128 and stack <= 16
    regs, stack = _find(usage, "hist_kernel")
    assert regs <= 40 and stack == 0


@pytest.mark.parametrize("model", ["9ModelUser", "10ModelTest4", "12ModelPeaks10", "11ModelStiff4"])
def test_shared_jacobian_kernel_fits_its_launch(usage, model):
    # jac0_kernel is launched <<<1, 256>>> (model_registry.cuh): registers x threads must fit one SM's
    # register file whatever the plug-in's row_model needs (round-1 advisor finding: 72 x 1024 did not)
    regs, stack = _find(usage, "jac0_kernelINS_" + model)
    assert regs * 256 <= 65536
