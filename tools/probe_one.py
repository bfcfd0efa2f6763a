"""Run one `depth base` through the C ABI on a given (or generated) BAM: target for ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sambamba_b200 as sb
import helpers
path = sys.argv[1] if len(sys.argv) > 1 else "/tmp/probe/one.bam"
if not os.path.exists(path):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    n = sys.argv[2] if len(sys.argv) > 2 else "200000"
    helpers.gen_bam(path, "-r", "chr20:64444167", "-n", n, "-s", "20", "-t", "32")
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
with sb.BDepth(path) as b:
    for _ in range(reps):
        b.run_base(collect=False)
    print(b.stats())
