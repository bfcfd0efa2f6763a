"""GPU parity tests: every call goes through the C ABI (libbdepth.so); the oracle is only the checker."""
import os

import numpy as np
import pytest

import helpers
from helpers import GOLDEN

pytestmark = pytest.mark.gpu

FIXTURES = ["issue_193.bam", "issue225.bam", "issue_204.bam", "mate_overlaps_1_3M_4M.bam"]


@pytest.fixture(scope="module")
def sb():
    import sambamba_b200
    return sambamba_b200


@pytest.fixture(scope="module")
def synth(tmp_path_factory):
    d = tmp_path_factory.mktemp("synth")
    return {
        "tiny": helpers.gen_bam(str(d / "tiny.bam"), "--preset", "tiny", "-t", 4),
        "mix": helpers.gen_bam(str(d / "mix.bam"), "--preset", "tiny", "-n", 20000, "--stored-every", 5, "-t", 4),
        "mid": helpers.gen_bam(str(d / "mid.bam"), "-r", "chrA:3000000", "-r", "chrB:500", "-r", "chrC:1000000", "-n", 400000, "-s", 7, "-t", 8),
    }


def all_paths(synth):
    return [os.path.join(GOLDEN, f) for f in FIXTURES] + list(synth.values())


def test_k1_inflate_bit_exact(sb, synth):
    for p in all_paths(synth):
        want = helpers.oracle_inflate(p)
        with sb.BDepth(p) as b:
            got = b.inflate()
        assert got.shape == want.shape and np.array_equal(got, want), p


def test_k2_scan_matches_record_walk(sb, synth):
    for p in all_paths(synth):
        u = helpers.oracle_inflate(p)
        first, refs = helpers.header_first_record_offset(u)
        recs = helpers.parse_records(u, first)
        with sb.BDepth(p) as b:
            assert b.refs == refs
            n, cols = b.scan(len(recs) + 10)
            st = b.stats()
        assert n == len(recs), p
        assert st["chain_fixups"] == 0
        off = np.array([r[0] for r in recs], np.uint64)
        assert np.array_equal(cols["rec_off"], off), p
        for name, idx in (("flag", 3), ("mapq", 4), ("n_cigar", 5)):
            assert np.array_equal(cols[name].astype(np.int64), np.array([r[idx] for r in recs], np.int64)), (p, name)
        placed = np.array([r[1] >= 0 and r[2] >= 0 for r in recs])
        assert np.array_equal(cols["ref_id"][placed], np.array([r[1] for r in recs], np.int32)[placed]), p
        assert np.array_equal(cols["pos"][placed], np.array([r[2] for r in recs], np.int32)[placed]), p


@pytest.mark.parametrize("minq", [0, 20])
def test_k3_base_counts_bit_exact(sb, synth, minq):
    for p in all_paths(synth):
        win = helpers.interesting_window(p)          # fixtures with a human-sized header: compare around the reads
        want, ost = helpers.oracle_counts(p, min_bq=minq, window=win)
        with sb.BDepth(p) as b:
            b.set_min_baseq(minq)
            got = b.run_base(window=win)
            st = b.stats()
        assert got.shape == want.shape, p
        bad = np.argwhere(got != want)
        assert bad.size == 0, (p, bad[:5], got[:, bad[0][1]] if bad.size else None, want[:, bad[0][1]] if bad.size else None)
        # a window query stages only the BAI chunks of the window: records outside them (unplaced reads) are never read
        assert st["n_records"] <= ost.n_records and st["n_records_pass"] == ost.n_pass, p
        assert st["covered_positions"] == int((want.sum(axis=0) > 0).sum()), p      # everything covered lies inside the window


def test_multi_batch_equals_single_batch(sb, synth):
    p = synth["mid"]
    want, _ = helpers.oracle_counts(p)
    with sb.BDepth(p) as b:
        b.set_tuning(batch_bytes=3 << 20)
        got = b.run_base()
        st = b.stats()
    assert st["n_batches"] > 5
    assert np.array_equal(got, want)


@pytest.mark.parametrize("chunk_blocks", [1, 3, 16])
def test_sub_batches_equal_single_pass(sb, synth, chunk_blocks):
    # every H2D chunk is a sub-batch whose scan/coverage runs while later chunks are still being inflated; records that
    # straddle a sub-batch boundary are carried in place.  Tiny chunks put a boundary after (almost) every BGZF block.
    p = synth["mid"]
    want, _ = helpers.oracle_counts(p)
    with sb.BDepth(p) as b:
        b.set_tuning(chunk_blocks=chunk_blocks)
        got = b.run_base()
        assert np.array_equal(got, want)
        b.set_tuning(batch_bytes=5 << 20, chunk_blocks=chunk_blocks)     # several batches of several sub-batches
        got = b.run_base()
        assert b.stats()["n_batches"] > 3
    assert np.array_equal(got, want)


def test_filter_none(sb, synth):
    p = synth["tiny"]
    want, _ = helpers.oracle_counts(p, mapq_gt=-1, flag_reject=0)
    with sb.BDepth(p) as b:
        b.set_filter(-1, 0)
        got = b.run_base()
    assert np.array_equal(got, want)
