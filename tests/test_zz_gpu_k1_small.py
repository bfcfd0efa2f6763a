"""The 4-warp K1 variant (BDEPTH_K1_STREAM_WARPS=4, kernels.cuh `k1_inflate_small`): same lane logic, other CTA shape.
An experiment for the end-to-end drain time that has not been on hardware yet: non-gating until seen green once."""
import os
import subprocess

import pytest

import helpers
from helpers import GOLDEN

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600), pytest.mark.xfail(strict=False, reason="first hardware run pending")]


def test_small_cta_inflate_gives_identical_output(tmp_path):
    p = helpers.gen_bam(str(tmp_path / "t.bam"), "-r", "chrA:900000", "-r", "chrB:600000", "-n", 120000, "-s", 4, "-t", 4, "--stored-every", 7)
    env = dict(os.environ, BDEPTH_K1_STREAM_WARPS="4")
    for args in (["base", os.path.join(GOLDEN, "issue_193.bam")], ["base", "-c", "0", p], ["window", "-w", "1000", "-T", "5", p]):
        a = subprocess.run([helpers.CLI] + args, capture_output=True)
        b = subprocess.run([helpers.CLI] + args, capture_output=True, env=env)
        assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
        assert a.stdout == b.stdout and len(a.stdout) > 100
