"""Kernel variants that are switched on by environment variables and have not been on hardware yet (non-gating until
seen green once): the 4-warp K1 CTA shape for streaming sub-launches (BDEPTH_K1_STREAM_WARPS=4, `k1_inflate_small`),
K1 with up to three literals per iteration (BDEPTH_K1_LIT3=1, `k1_inflate_lit3`) and k3_gather with lane-parallel record
prefetch (BDEPTH_K3_PREFETCH=1).  Both must print what the default kernels print."""
import os
import subprocess

import pytest

import helpers
from helpers import GOLDEN

N_READS = 12000 if os.environ.get("BDEPTH_EMULATE") == "1" else 120000

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def test_small_cta_inflate_gives_identical_output(tmp_path):
    p = helpers.gen_bam(str(tmp_path / "t.bam"), "-r", "chrA:900000", "-r", "chrB:600000", "-n", N_READS, "-s", 4, "-t", 4, "--stored-every", 7)
    _same_output_with(tmp_path, p, dict(BDEPTH_K1_STREAM_WARPS="4"))


def test_three_literal_inflate_gives_identical_output(tmp_path):
    p = helpers.gen_bam(str(tmp_path / "t.bam"), "-r", "chrA:900000", "-r", "chrB:600000", "-n", N_READS, "-s", 6, "-t", 4, "--stored-every", 5)
    _same_output_with(tmp_path, p, dict(BDEPTH_K1_LIT3="1"))


def test_k3_prefetch_gives_identical_output(tmp_path):
    p = helpers.gen_bam(str(tmp_path / "t.bam"), "-r", "chrA:900000", "-r", "chrB:600000", "-n", N_READS, "-s", 5, "-t", 4)
    _same_output_with(tmp_path, p, dict(BDEPTH_K3_PREFETCH="1"), extra=[["base", "-q", "25", p], ["base", "-c", "0", os.path.join(GOLDEN, "issue225.bam")]])


def _same_output_with(tmp_path, p, envadd, extra=()):
    env = dict(os.environ, **envadd)
    for args in list(extra) + [["base", os.path.join(GOLDEN, "issue_193.bam")], ["base", "-c", "0", p], ["window", "-w", "1000", "-T", "5", p]]:
        a = subprocess.run([helpers.CLI] + args, capture_output=True)
        b = subprocess.run([helpers.CLI] + args, capture_output=True, env=env)
        assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
        assert a.stdout == b.stdout and len(a.stdout) > 100
