"""Parity at BASELINE configs[1] size (synthetic 30x chr20: 2.26 GB BAM, 3.76 GB inflated, 12.9 M reads) through
properties that do not need the oracle to finish a full run: the file's own BGZF CRC32s (checksum of checksums) for K1,
coordinate order and column sanity for K2, the span / counter identity and the covered-position count for K3, and
batch-invariance of the whole pass.  The checkers are judged on the CPU in tests/test_fullsize_props_cpu.py.
Gating since round 2 (seen green on hardware in round 1)."""
import os
import sys

import numpy as np
import pytest

import fullsize_props as fp
import helpers

EMULATE = os.environ.get("BDEPTH_EMULATE") == "1"      # the test's own logic is checked on the CPU at small size (tests/conftest.py)

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1500),
              pytest.mark.skipif(os.environ.get("BDEPTH_SKIP_FULLSIZE") == "1" and not EMULATE, reason="BDEPTH_SKIP_FULLSIZE=1 (development runs only: the full-size parity tests gate the suite by default)")]


@pytest.fixture(scope="module")
def chr20(tmp_path_factory):
    if EMULATE:
        return helpers.gen_bam(str(tmp_path_factory.mktemp("fs") / "small.bam"), "-r", "chr20:300000", "-n", 60000, "-s", 20, "-t", 4)
    sys.path.insert(0, helpers.ROOT)
    import bench
    return bench.ensure_workload(1, bench.READS_PER_UNIT)       # the bench workload; generated once per box (about a minute)


def test_full_size_properties(chr20):
    import sambamba_b200 as sb
    raw = np.fromfile(chr20, dtype=np.uint8)
    with sb.BDepth(chr20) as b:
        n_ref = len(b.refs)
        u = b.inflate()
        assert fp.inflate_matches_the_files_own_checksums(memoryview(raw), u), "K1: inflated bytes do not match the CRC32s stored in the file"
        del u
        n, cols = b.scan(14_000_000)
        assert (n == 60000 if EMULATE else 12_000_000 < n <= 14_000_000) and all(len(v) == n for v in cols.values())
        assert fp.scan_is_sorted_and_consistent(cols, n_ref), "K2: records out of coordinate order or inconsistent columns"
        counts = b.run_base()
        st = b.stats()
        assert st["n_records"] == n and st["n_records_pass"] == int(fp.passing(cols).sum())
        assert fp.counters_add_up(cols, counts, st["covered_positions"]), "K3: counters do not add up to the reads' reference spans"
        # the same pass in 512 MB batches (16 instead of 1) and with the input staged in HBM first: identical counters
        b.set_tuning(1 << 20 if EMULATE else 512 << 20, 0)
        again = b.run_base()
        assert b.stats()["n_batches"] > 4 and np.array_equal(again, counts)
        del again
        b.set_tuning(6 << 30, 0)
        b.stage()
        assert np.array_equal(b.run_base(), counts)


def test_full_size_bit_exact_against_the_oracle(chr20):
    """BASELINE configs[1] ("bit-exact vs CPU") at the size the headline is quoted on: every one of the 7 x 64.4 M
    counters of `depth base` from the GPU equals the oracle's (closed-form scatter, checked against the faithful
    column sweep on every fixture in tests/test_oracle_golden.py), streamed and with the input staged."""
    import sambamba_b200 as sb
    want, ost = helpers.oracle_counts(chr20, threads=min(64, os.cpu_count() or 8))
    with sb.BDepth(chr20) as b:
        got = b.run_base()
        st = b.stats()
        assert got.shape == want.shape
        assert np.array_equal(got, want), f"GPU counters differ from the oracle at {int(np.argmax((got != want).any(axis=0)))}"
        assert st["n_records"] == ost.n_records and st["n_records_pass"] == ost.n_pass and st["covered_positions"] == ost.covered
        del got
        b.stage()
        assert np.array_equal(b.run_base(), want)


@pytest.mark.skipif(EMULATE, reason="the CLI comparison at small size is tests/test_gpu_cli.py")
def test_full_size_cli_output_is_identical_to_the_oracle_cli(chr20, tmp_path):
    """`depth base` through the drop-in CLI (rows formatted on the GPU) against the oracle CLI: 2.15 GB of text, compared by digest."""
    import hashlib
    import subprocess

    def digest(cmd):
        p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
        h, n = hashlib.md5(), 0
        for chunk in iter(lambda: p.stdout.read(1 << 24), b""):
            h.update(chunk)
            n += len(chunk)
        assert p.wait() == 0, cmd
        return h.hexdigest(), n
    got = digest([helpers.CLI, "base", chr20])
    want = digest([helpers.ORACLE_EXE, "depth", "base", chr20])
    assert got == want and got[1] > 1_000_000_000
