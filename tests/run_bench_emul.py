"""TEST INFRASTRUCTURE: bench.py's own arm with the product library replaced by its CUDA-on-CPU emulation build (tests/emul/libbdepth_emul.so),
so that the bench's control flow -- warm-up, timed passes, A/B legs, e2e, text rows, verification under its deadline, the JSON line -- is exercised
without a GPU.  The numbers it prints are meaningless.    python tests/run_bench_emul.py --reads-per-unit 20000 --steps 1 --warmup 1"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sambamba_b200._lib as L      # noqa: E402

L.lib_path = lambda: os.path.join(ROOT, "tests", "emul", "libbdepth_emul.so")
L._lib = None
import bench                        # noqa: E402

if __name__ == "__main__":
    sys.exit(bench.main())
