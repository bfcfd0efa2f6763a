"""Every selectable kernel variant must print what the default kernels print: the round-1 one-phase inflater for all blocks
(BDEPTH_K1_ONEPHASE=1; by default it only takes the header blocks and what phase 1 hands back), the phase-1 instantiations of the
two-phase inflater (BDEPTH_K1H_VARIANT: limits in registers / in shared memory, 4 / 5 / 6 CTAs per SM; by default chosen by launch
size), round 1's k3_gather with and without its record prefetch (BDEPTH_K3=gather, BDEPTH_K3_PREFETCH=0; default: k3_tile), the byte-flattened
phase 2 (BDEPTH_K1LZ=flat)."""
import os
import subprocess

import pytest

import helpers
from helpers import GOLDEN

N_READS = 12000 if os.environ.get("BDEPTH_EMULATE") == "1" else 120000

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.fixture(scope="module")
def bam(tmp_path_factory):
    d = tmp_path_factory.mktemp("variants")
    return helpers.gen_bam(str(d / "t.bam"), "-r", "chrA:900000", "-r", "chrB:600000", "-n", N_READS, "-s", 4, "-t", 4, "--stored-every", 7)


@pytest.mark.parametrize("envadd", [dict(BDEPTH_K1_ONEPHASE="1"), dict(BDEPTH_K1H_VARIANT="0"), dict(BDEPTH_K1H_VARIANT="1"), dict(BDEPTH_K1H_VARIANT="2"),
                                    dict(BDEPTH_K1H_VARIANT="3"), dict(BDEPTH_K3="gather"), dict(BDEPTH_K3="gather", BDEPTH_K3_PREFETCH="0"), dict(BDEPTH_K1LZ="flat"), dict(BDEPTH_K1LZ="v12")],
                         ids=lambda e: ",".join(f"{k[7:]}={v}" for k, v in e.items()))
def test_variant_gives_identical_output(bam, envadd):
    env = dict(os.environ, **envadd)
    for args in (["base", os.path.join(GOLDEN, "issue_193.bam")], ["base", "-c", "0", bam], ["base", "-q", "25", bam], ["window", "-w", "1000", "-T", "5", bam],
                 ["base", "-c", "0", os.path.join(GOLDEN, "issue225.bam")]):
        a = subprocess.run([helpers.CLI] + args, capture_output=True)
        b = subprocess.run([helpers.CLI] + args, capture_output=True, env=env)
        assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
        assert a.stdout == b.stdout and len(a.stdout) > 100
