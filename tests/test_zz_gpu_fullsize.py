"""Parity at BASELINE configs[1] size (synthetic 30x chr20: 2.26 GB BAM, 3.76 GB inflated, 12.9 M reads) through
properties that do not need the oracle to finish a full run: the file's own BGZF CRC32s (checksum of checksums) for K1,
coordinate order and column sanity for K2, the span / counter identity and the covered-position count for K3, and
batch-invariance of the whole pass.  The checkers are judged on the CPU in tests/test_fullsize_props_cpu.py.
Written after the round's GPU budget was spent: non-gating until seen green once (then the xfail marker goes)."""
import os
import sys

import numpy as np
import pytest

import fullsize_props as fp
import helpers

EMULATE = os.environ.get("BDEPTH_EMULATE") == "1"      # the test's own logic is checked on the CPU at small size (tests/conftest.py)

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1500),
              pytest.mark.skipif(os.environ.get("BDEPTH_FULLSIZE") != "1" and not EMULATE, reason="several minutes (generates and walks a 2.3 GB BAM): set BDEPTH_FULLSIZE=1, as tools/gpu_round_start.sh does"),
              pytest.mark.xfail(strict=False, reason="first hardware run pending (written after the round's GPU budget was spent)")]


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
