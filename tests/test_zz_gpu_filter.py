"""`-F` queries through the GPU (k2_decode<true> + filter.cuh): the CLI with `-F query` on a file must print what the
oracle prints with `-F ""` on the same file reduced to the reads a Python statement of the query keeps.

Compiler and evaluator are the code tests/test_emul_filter.py runs on the CPU; only the kernel instantiation and the
program upload are new on hardware.  Written after the round's GPU budget was spent: non-gating (xfail, non-strict)
until seen green once."""
import os

import pytest

import helpers
import test_emul_filter as tef

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def test_cli_filters_match_oracle_on_prefiltered_input(tmp_path):
    p = tef.make_bam(str(tmp_path / "f.bam"), seed=5, n=3000, empty_seq=False)
    u = helpers.oracle_inflate(p)
    _, recs = tef.parse_all(u)
    full = os.environ.get("BDEPTH_FULLSIZE") == "1"
    modes = (["base", "-c", "0"], ["window", "-w", "500", "-T", "2"], ["region", "-L", "c1:100-3000", "-T", "1"])
    for k, (q, fn) in enumerate(tef.QUERIES):
        if not full and k % 3 != 1:
            continue                      # every process start pays a CUDA context: a third of the queries, one mode each, unless asked for all
        sub = helpers.subset_bam(p, str(tmp_path / f"sub{k}.bam"), [bool(fn(r)) for r in recs])
        for mode in (modes if full else modes[k % 3:k % 3 + 1]):
            rc1, out1, err1 = helpers.run_cli(mode + ["-F", q, p])
            rc2, out2, err2 = helpers.oracle_cli(mode + ["-F", "", sub])
            assert rc1 == 0 and rc2 == 0, (q, err1, err2)
            assert out1 == out2, (q, mode)


def test_cli_refuses_what_it_cannot_evaluate(tmp_path):
    p = tef.make_bam(str(tmp_path / "f.bam"), seed=6, n=200)
    for q in tef.BAD[::1 if os.environ.get("BDEPTH_FULLSIZE") == "1" else 4]:
        if q == "":
            continue
        rc, out, err = helpers.run_cli(["base", "-F", q, p])
        assert rc == 1 and err.startswith(b"sambamba-depth: "), q


def test_filter_with_fix_mates(tmp_path):
    import test_emul_mates as tem
    p = tem.make_pairs_bam(str(tmp_path / "pm.bam"), 21)
    u = helpers.oracle_inflate(p)
    _, recs = tef.parse_all(u)
    sub = helpers.subset_bam(p, str(tmp_path / "sub.bam"), [r.mapq >= 20 and not r.flag & 0x400 for r in recs])
    rc1, out1, err1 = helpers.run_cli(["base", "-m", "-c", "0", "-F", "mapping_quality >= 20 and not duplicate", p])
    rc2, out2, _ = helpers.oracle_cli(["base", "-m", "-c", "0", "-F", "", sub])
    assert rc1 == 0 and rc2 == 0, err1
    assert out1 == out2
