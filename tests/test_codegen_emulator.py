"""Code-generator check without a GPU: the product's host sources (graph, scheduler, ring/register allocation) are
linked against the tape-ISA emulator of tests/emu/ and driven through the same C ABI by the GPU parity tests.
The emulator rejects every tape that arms, waits for, reads or re-arms a TMA ring slot illegally, that reads back a
buffer the same launch wrote, or that reads a register-file slot before writing it, and it computes each opcode with
its documented meaning, so value mismatches against the oracle show scheduling errors too.
(Brownian/MT19937 kernels are not tape-driven and are covered on the GPU only.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libfmcuda_emu.so")


@pytest.fixture(scope="module")
def emulator():
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    assert os.path.exists(EMU_LIB)
    return EMU_LIB


def run_parity(emulator, extra_env=None, select="not brownian and not mt19937 and not pool_recycles"):
    env = dict(os.environ, FMC_TEST_TAPE_EMULATOR=emulator)
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider",
           os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-k", select]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]


def test_parity_suite_on_the_emulator(emulator):
    run_parity(emulator)


@pytest.mark.parametrize("opts", ["pipeline=0", "ring_max=2,ring_min=1", "ring_max=3,horizon=4", "ring_max=16,target_ctas=1",
                                  "grid_limit=1", "grid_limit=3,max_sets=2", "grid_limit=2,max_sets=3,ring_max=2,ring_min=1"])
def test_parity_suite_on_the_emulator_with_scheduler_knobs(emulator, opts):
    run_parity(emulator, {"FMC_TEST_OPTIONS": opts}, select="compound or ragged or reductions or fused or long_tape or unfused")


def test_windows_of_time_steps_on_the_emulator(emulator):
    """Option window_levels: the LMM simulation emitted in windows of several time steps (component-major, running sums kept in the
    register file, T_RATIOACC_A) gives bit-identical LIBORs to one launch per step, in every geometry; the emulator also checks the
    ring / register-file discipline of the long window tapes."""
    env = dict(os.environ, LD_PRELOAD=emulator, FMC_EMU_FAKE_BROWNIAN="1")
    r = subprocess.run([sys.executable, os.path.join(EMU_DIR, "run_window_check.py"), "700"], env=env, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert r.stdout.count("LIBORs bit-identical True") >= 6, r.stdout


@pytest.mark.parametrize("env", [{}, {"FMC_HOST_THREADS": "4", "FMC_HOST_SPIN_US": "0"}, {"FMC_HOST_THREADS": "3", "FMC_HOST_SPIN_US": "300", "FMC_HOST_BLOCKS_PER_THREAD": "7"},
                                 {"FMC_HOST_AVX": "0", "FMC_HOST_THREADS": "1"}])
def test_host_staging_pool_on_the_emulator(emulator, env):
    """createRandomVariable(time, double[]) -> getRealizations() through the host worker pool (csrc/runtime.cpp: HostWorkers): ragged sizes,
    special values, both conversion loops, spinning and sleeping workers."""
    e = dict(os.environ, LD_PRELOAD=emulator, **env)
    r = subprocess.run([sys.executable, os.path.join(EMU_DIR, "run_upload_check.py")], env=e, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "round trips ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_launch_count_of_a_repeated_lmm_step_is_stable(emulator):
    """The objective function of the calibration records the same graph thousands of times: every repetition must be cut into the same
    windows (Runtime::run_windows counts each still-referenced target once, however often its recycled node slot is listed)."""
    env = dict(os.environ, LD_PRELOAD=emulator, FMC_EMU_FAKE_BROWNIAN="1")
    r = subprocess.run([sys.executable, os.path.join(EMU_DIR, "run_launch_count.py"), "1024"], env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "launch count stable True" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
