"""The `gpu` tests of the CLI, of `-m`, of `-F`, of the BAI builder and of the multi-rank sharding (ranks as threads over an NCCL stand-in), run on the CPU against tests/emul/libbdepth_emul.so: the same host
pipeline (bdepth.cu) and kernels compiled with g++ over a CUDA-on-CPU emulation (tests/emul/cuda_shim.hpp: a fiber per
thread, rendezvous for warp collectives and __syncthreads).  TEST INFRASTRUCTURE: it shows that launch plumbing written
without access to a GPU is logically right; it is no substitute for the hardware run (memory model, alignment, PTX paths,
speed) and no part of the product.  The other gpu suites pass under it as well (parity, edge cases, sparse staging,
kernel variants: about 15 minutes, run them with BDEPTH_EMULATE=1 python -m pytest tests -m gpu --runxfail)."""
import os
import subprocess
import sys

from helpers import ROOT

# (suite, -k expression): the -m suite is split in two so that the four processes take about the same time
SUITES = [("tests/test_gpu_cli.py", None), ("tests/test_zz_gpu_mates.py", "several_batches or window_mode"),
          ("tests/test_zz_gpu_mates.py", "not several_batches and not window_mode"), ("tests/test_zz_gpu_filter.py", None), ("tests/test_gpu_multi.py", None), ("tests/test_zzz_gpu_index.py", None), ("tests/test_zzy_gpu_fuzz_regressions.py", None)]


def test_the_emulation_itself():
    """Known-answer kernels for every collective the library uses (ballot with shrinking masks, shuffles, reductions,
    __syncthreads with threads that left, static / dynamic shared memory, atomics): tests/emul/shim_selftest.cpp."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emul"), "shim_selftest"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(ROOT, "tests", "emul", "shim_selftest")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr


def test_gpu_suites_pass_under_cpu_emulation():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emul")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    env = dict(os.environ, BDEPTH_EMULATE="1")
    procs = [((s, k), subprocess.Popen([sys.executable, "-m", "pytest", s, "-q", "-m", "gpu", "--runxfail", "-p", "no:cacheprovider", "-x"] + (["-k", k] if k else []), cwd=ROOT, env=env,
                                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)) for s, k in SUITES]
    for s, p in procs:
        out, _ = p.communicate(timeout=1500)
        assert p.returncode == 0, (s, out[-3000:])
        assert " passed" in out and "failed" not in out and "skipped" not in out, (s, out[-600:])      # (deselected by -k is fine)
