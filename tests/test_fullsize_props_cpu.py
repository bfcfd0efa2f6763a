"""The property checkers of tests/fullsize_props.py judged on the CPU at small size: fed with the oracle's outputs they
must accept, fed with corrupted ones they must refuse."""
import numpy as np

import fullsize_props as fp
import helpers


def _cols(u, refs_len):
    first, refs = helpers.header_first_record_offset(u)
    rows = helpers.parse_records(u, first)
    n = len(rows)
    cols = dict(ref_id=np.array([r[1] for r in rows], np.int32), pos=np.array([r[2] for r in rows], np.int32), flag=np.array([r[3] for r in rows], np.uint16),
                mapq=np.array([r[4] for r in rows], np.uint8), n_cigar=np.array([r[5] for r in rows], np.uint16), rec_off=np.array([r[0] for r in rows], np.uint64))
    span = np.array([r[6] for r in rows], np.int64)
    room = np.array([max(0, refs[r[1]][1] - r[2]) if 0 <= r[1] < len(refs) else 0 for r in rows], np.int64)
    cols["span"] = np.minimum(span, room).astype(np.uint32)       # the scan reports the span clipped to the reference
    return cols, refs, n


def test_checkers_accept_the_oracle_and_refuse_corruption(tmp_path):
    p = helpers.gen_bam(str(tmp_path / "t.bam"), "--preset", "tiny", "-n", 8000, "-t", 2, "--stored-every", 9)
    raw = open(p, "rb").read()
    u = helpers.oracle_inflate(p)
    assert fp.inflate_matches_the_files_own_checksums(raw, u)
    bad = u.copy()
    bad[len(bad) // 2] ^= 1
    assert not fp.inflate_matches_the_files_own_checksums(raw, bad)
    cols, refs, n = _cols(u, None)
    assert fp.scan_is_sorted_and_consistent(cols, len(refs))
    swapped = {k: v.copy() for k, v in cols.items()}
    swapped["pos"][[10, 2000]] = swapped["pos"][[2000, 10]]
    assert not fp.scan_is_sorted_and_consistent(swapped, len(refs))
    counts, st = helpers.oracle_counts(p)
    assert int(fp.passing(cols).sum()) == st.n_pass
    assert fp.counters_add_up(cols, counts, st.covered)
    c2 = counts.copy()
    c2[3, 12345] += 1
    assert not fp.counters_add_up(cols, c2, st.covered)
