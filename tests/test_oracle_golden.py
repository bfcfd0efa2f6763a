"""Pins the CPU oracle (oracle/depth_oracle.c) on the reference's own golden vectors:
test/test_suite.sh:149-154 (issue_193), :156-162 (issue_204), :176-194 (issue225), and the column
known-answer test in BioD/bio/std/hts/bam/pileup.d:699-857.  No GPU needed."""
import os
import struct
import zlib

import numpy as np
import pytest

import helpers
from helpers import GOLDEN


def G(f):
    return os.path.join(GOLDEN, f)


def test_issue_193_depth_base():
    rc, out, err = helpers.oracle_cli(["base", G("issue_193.bam")])
    assert rc == 0
    assert out == open(G("issue_193_expected_output.txt"), "rb").read()
    assert b"Processing reference #1" in err


def test_issue_225_depth_base_with_and_without_L():
    for extra in ([], ["-L", "chrM"]):
        rc, out, _ = helpers.oracle_cli(["base", "-c", "1"] + extra + [G("issue225.bam")])
        assert rc == 0 and out == open(G("issue225.out"), "rb").read(), extra
        rc, out, _ = helpers.oracle_cli(["base", "-c", "0"] + extra + [G("issue225.bam")])
        assert rc == 0 and out == open(G("issue225.z.out"), "rb").read(), extra


def test_issue_204_region_fix_mate_overlaps():
    rc, out, _ = helpers.oracle_cli(["region", G("issue_204.bam"), "-L", "2:166868600-166868813", "-T", "15", "-T", "20", "-T", "25", "-m"])
    assert rc == 0
    assert out == open(G("issue_204_expected_output.txt"), "rb").read()


def test_closed_form_equals_sweep_on_all_fixtures(tmp_path):
    """The per-read scatter oracle (used at full size) must agree with the column sweep (pinned above)."""
    files = [G(f) for f in ("issue_193.bam", "issue225.bam", "issue_204.bam", "mate_overlaps_1_3M_4M.bam")]
    files.append(helpers.gen_bam(str(tmp_path / "t.bam"), "--preset", "tiny", "-t", 2))
    for p in files:
        for minq in (0, 25):
            win = helpers.interesting_window(p)
            counts, _ = helpers.oracle_counts(p, min_bq=minq, window=win)
            rc, out, _ = helpers.oracle_cli(["base", "-q", str(minq), p])
            assert rc == 0
            # rows: REF POS COV A C G T DEL REFSKIP SAMPLE ; N = COV - (A+C+G+T+DEL+REFSKIP)
            with helpers_refs(p) as lin0:
                seen = 0
                for line in out.splitlines()[1:]:
                    f = line.split(b"\t")
                    g = lin0[f[0].decode()] + int(f[1]) - win[0]
                    a, c, gg, t, d, s = (int(x) for x in f[3:9])
                    cov = int(f[2])
                    col = counts[:, g]
                    assert (col[0], col[1], col[2], col[3], col[5], col[6]) == (a, c, gg, t, d, s), (p, line)
                    assert int(col.sum()) == cov, (p, line)
                    seen += 1
                assert seen == int((counts.sum(axis=0) > 0).sum()), p


def test_closed_form_for_fix_mate_overlaps_equals_sweep(tmp_path):
    """`-m` in base mode as a per-name-group closed form (groundwork for the GPU kernel) against the faithful sweep."""
    import random
    r = random.Random(8)
    files = [G("issue_204.bam"), G("mate_overlaps_1_3M_4M.bam")]
    # hand-made pairs: overlapping mates with deletions / skips / different MAPQ, three reads of one name, same name on
    # two references, a filtered mate
    L = 4000
    reads = []
    def seq(n):
        return "".join(r.choice("ACGT") for _ in range(n))
    for i in range(120):
        s1 = r.randrange(0, L - 400); s2 = s1 + r.randrange(0, 140)
        c1 = r.choice([[(100, 0)], [(40, 0), (5, 2), (55, 0)], [(30, 0), (20, 3), (60, 0)], [(10, 4), (90, 0)], [(50, 0), (4, 1), (46, 0)]])
        c2 = r.choice([[(100, 0)], [(20, 0), (8, 2), (72, 0)], [(60, 0), (30, 3), (30, 0)], [(95, 0), (5, 4)]])
        q1 = sum(l for l, op in c1 if op in (0, 1, 4)); q2 = sum(l for l, op in c2 if op in (0, 1, 4))
        reads.append((0, s1, r.choice([60, 60, 20, 7]), 0x63, c1, seq(q1), f"p{i}"))
        reads.append((0, s2, r.choice([60, 60, 20, 7, 0]), 0x93, c2, seq(q2), f"p{i}"))
        if i % 17 == 0:
            reads.append((0, s1 + 20, 50, 0x800, [(80, 0)], seq(80), f"p{i}"))          # a third read of the same name
        if i % 23 == 0:
            reads.append((1, 100 + i, 60, 0, [(50, 0)], seq(50), f"p{i}"))              # same name on another reference
    reads.sort(key=lambda x: (x[0], x[1]))
    quals = [[r.randint(2, 41) for _ in x[5]] for x in reads]
    files.append(helpers.write_bam(str(tmp_path / "pairs.bam"), [("c1", L), ("c2", 600)], reads, quals=quals))
    total_fixes = 0
    for p in files:
        for minq in (0, 20):
            win = helpers.interesting_window(p)
            counts, fixes = helpers.oracle_counts_fix_mates(p, min_bq=minq, window=win)
            plain, _ = helpers.oracle_counts(p, min_bq=minq, window=win)
            total_fixes += fixes
            assert fixes > 0 and (counts != plain).any(), p
            # default -c 1: only covered positions are printed (the fixtures carry human-genome headers)
            rc, out, _ = helpers.oracle_cli(["base", "-m", "-q", str(minq), p])
            assert rc == 0
            seen = 0
            with helpers_refs(p) as lin0:
                for line in out.splitlines()[1:]:
                    f = line.split(b"\t")
                    g = lin0[f[0].decode()] + int(f[1]) - win[0]
                    a, c, gg, t, d, s = (int(x) for x in f[3:9])
                    col = counts[:, g]
                    assert (col[0], col[1], col[2], col[3], col[5], col[6]) == (a, c, gg, t, d, s), (p, minq, line, col)
                    assert int(col.sum()) == int(f[2]), (p, minq, line)
                    seen += 1
            assert seen == int((counts.sum(axis=0) > 0).sum()), p
    assert total_fixes > 1000


class helpers_refs:
    def __init__(self, path):
        u = helpers.oracle_inflate(path)
        _, refs = helpers.header_first_record_offset(u)
        self.lin0, lin = {}, 0
        for name, L in refs:
            self.lin0[name] = lin
            lin += L

    def __enter__(self):
        return self.lin0

    def __exit__(self, *a):
        return False


# ---- the pileup.d unittest reads (BioD/bio/std/hts/bam/pileup.d:703-740) written as a BAM by hand
SEQS = ["ATTATGGACATTGTTTCCGTTATCATCATCATCATCATCATCATCATTATCATC", "GACATTGTTTCCGTTATCATCATCATCATCATCATCATCATCATCATCATCATC",
        "ATTGTTTCCGTTATCATCATCATCATCATCATCATCATCATCATCATCATCACC", "TGTTTCCGTTATCATCATCATCATCATCATCATCATCATCATCATCATCACCAC",
        "TCCGTTATCATCATCATCATCATCATCATCATCATCATCATCATCACCACCACC", "GTTATCATCATCATCATCATCATCATCATCATCATCATCATCATCGTCACCCTG",
        "TCATCATCATCATAATCATCATCATCATCATCATCATCGTCACCCTGTGTTGAG", "TCATCATCATCGTCACCCTGTGTTGAGGACAGAAGTAATTTCCCTTTCTTGGCT",
        "TCATCATCATCATCACCACCACCACCCTGTGTTGAGGACAGAAGTAATATCCCT", "CACCACCACCCTGTGTTGAGGACAGAAGTAATTTCCCTTTCTTGGCTGGTCACC"]
CIGARS = [[(54, 0)], [(54, 0)], [(50, 0), (3, 1), (1, 0)], [(54, 0)], [(54, 0)], [(54, 0)], [(2, 4), (52, 0)],
          [(16, 0), (15, 2), (38, 0)], [(13, 0), (3, 1), (38, 0)], [(54, 0)]]
POS = [758, 764, 767, 769, 773, 776, 785, 795, 804, 817]


def test_pileup_unittest_columns(tmp_path):
    reads = [(0, POS[i], 60, 0, CIGARS[i], SEQS[i], f"r{i}") for i in range(10)]
    p = helpers.write_bam(str(tmp_path / "u.bam"), [("20", 2000)], reads)
    rc, out, _ = helpers.oracle_cli(["base", p])
    assert rc == 0
    rows = {int(l.split(b"\t")[1]): [int(x) for x in l.split(b"\t")[2:9]] for l in out.splitlines()[1:]}

    def bases(s):   # expected column string -> (COV, A, C, G, T, DEL, REFSKIP)
        return [len(s), s.count("A"), s.count("C"), s.count("G"), s.count("T"), s.count("-"), 0]
    # pileup.d:806-840: the column strings asserted by the reference's own unit test
    assert rows[796] == bases("CCCCCCAC")
    assert rows[805] == bases("TCCCCCCCC")
    assert rows[806] == bases("AAAAAAAGA")
    assert rows[821] == bases("AAGG-AA")
    assert rows[826] == bases("CCCCCC")
    assert rows[849] == bases("TAT")
    # and the closed form agrees
    counts, _ = helpers.oracle_counts(p)
    for pos, r in rows.items():
        col = counts[:, pos]
        assert [int(col.sum()), col[0], col[1], col[2], col[3], col[5], col[6]] == r


def _rows(out):
    return [l.split("\t") for l in out.decode().splitlines() if l and not l.startswith("#")]


def test_closed_form_segment_stats_equal_the_sweep(tmp_path):
    """oracle_segment_stats (closed form: bench.py verifies full-size window / region runs with it) against the faithful sweep's
    `depth window` and `depth region` output: readCount, meanCoverage (float32, %g) and the -T percentages, with and without -q."""
    refs = [("chrA", 40000), ("chrB", 900), ("chrC", 25000)]
    p = helpers.gen_bam(str(tmp_path / "s.bam"), "-r", "chrA:40000", "-r", "chrB:900", "-r", "chrC:25000", "-n", 9000, "-s", 5, "-t", 2)
    lin0 = np.concatenate([[0], np.cumsum([l for _, l in refs])])
    thr = [1, 5, 12]
    for q in (0, 25):
        qa = ["-q", str(q)] if q else []
        # windows: every full 1000-bp window of every reference (all have reads)
        rc, out, _ = helpers.oracle_cli(["window", "-w", "1000"] + sum((["-T", str(t)] for t in thr), []) + qa + [p])
        assert rc == 0
        rows = _rows(out)
        name_to_i = {n: i for i, (n, _) in enumerate(refs)}
        a = np.array([lin0[name_to_i[r[0]]] + int(r[1]) for r in rows], np.uint64)
        b = np.array([lin0[name_to_i[r[0]]] + int(r[2]) for r in rows], np.uint64)
        assert len(rows) == sum(l // 1000 for _, l in refs) and (np.diff(a.astype(np.int64)) > 0).all()
        reads, bases, cov = helpers.oracle_segment_stats(p, a, b, thr, min_bq=q, threads=3)
        for i, r in enumerate(rows):
            ln = np.float32(int(b[i] - a[i]))
            want = [str(int(reads[i])), "%g" % (np.float32(bases[i]) / ln)] + ["%g" % (np.float32(100) * np.float32(cov[t][i]) / ln) for t in range(len(thr))]
            assert r[3:3 + 2 + len(thr)] == want, (q, i, r, want)
        # regions from a BED file (sorted, disjoint)
        rnd = np.random.RandomState(3)
        bed = []
        for ri, (n, l) in enumerate(refs):
            x = 0
            while True:
                x += int(rnd.randint(1, 400))
                e = x + int(rnd.randint(1, 700))
                if e > l:
                    break
                bed.append((n, x, e))
                x = e
        bp = str(tmp_path / "r.bed")
        open(bp, "w").write("".join(f"{n}\t{s}\t{e}\n" for n, s, e in bed))
        rc, out, _ = helpers.oracle_cli(["region", "-L", bp] + sum((["-T", str(t)] for t in thr), []) + qa + [p])
        assert rc == 0
        rows = _rows(out)
        assert len(rows) == len(bed)
        a = np.array([lin0[name_to_i[n]] + s for n, s, e in bed], np.uint64)
        b = np.array([lin0[name_to_i[n]] + e for n, s, e in bed], np.uint64)
        reads, bases, cov = helpers.oracle_segment_stats(p, a, b, thr, min_bq=q, threads=2)
        for i, r in enumerate(rows):
            ln = np.float32(int(b[i] - a[i]))
            want = [str(int(reads[i])), "%g" % (np.float32(bases[i]) / ln)] + ["%g" % (np.float32(100) * np.float32(cov[t][i]) / ln) for t in range(len(thr))]
            assert r[3:3 + 2 + len(thr)] == want, (q, i, r, want)


@pytest.mark.parametrize("name,bytewise", [("issue225.bam", True), ("issue_193.bam", True), ("issue_204.bam", True), ("mate_overlaps_1_3M_4M.bam", False)])
def test_index_builder_reproduces_the_reference_bai_files(name, bytewise):
    """oracle_build_bai (IndexBuilder, BioD/bio/std/hts/bam/bai/indexing.d) is pinned on the four .bai files the reference ships next
    to its test BAMs -- sambamba's own indexer wrote them: bins, chunks, metadata pseudo-bins, linear index and n_no_coor are equal;
    three files are equal byte for byte, the fourth differs only in the order of its bins (the reference: iteration order of a D
    associative array; the oracle: ascending)."""
    p = os.path.join(GOLDEN, name)
    mine = helpers.oracle_build_bai(p, threads=2)
    ref = open(p + ".bai", "rb").read()
    assert helpers.parse_bai(mine) == helpers.parse_bai(ref)
    assert len(mine) == len(ref) and (mine == ref) == bytewise


def test_index_builder_edge_cases_by_hand(tmp_path):
    """A hand-checked case: chunk ends where the bin changes, position-less reads only count, gaps of the linear index are filled."""
    M = 0
    seq = "ACGT" * 10
    reads = [(0, 100, 60, 0, [(40, M)], seq, "a"), (0, 200, 60, 0, [(40, M)], seq, "b"), (0, 16380, 60, 0, [(40, M)], seq, "c"), (0, 50000, 60, 0, [(40, M)], seq, "d"),
             (-1, -1, 0, 4, [], seq, "u")]
    p = helpers.write_bam(str(tmp_path / "h.bam"), [("r0", 100000), ("r1", 500)], reads, bins="auto", index=False)
    (refs, no_coor) = helpers.parse_bai(helpers.oracle_build_bai(p))
    bins, lin = refs[0]
    assert no_coor == 1 and refs[1] == ({}, [])
    assert sorted(bins) == [585, 4681, 4684, 37450]      # a, b in leaf 4681; c crosses 16384 -> bin 585; d in leaf 4684
    assert bins[37450][1] == (4, 0)
    assert len(lin) == 4 and lin[0] == bins[4681][0][0] and lin[1] == bins[585][0][0] and lin[2] == lin[1] and lin[3] == bins[4684][0][0]      # window 2 is a gap: filled from the left
    assert bins[4681][0][1] == bins[585][0][0] and bins[585][0][1] == bins[4684][0][0]          # chunks meet where the bin changes


LEAD_N_READS = [  # (pos, cigar, l_seq): CIGARs that begin with N and do NOT end in M/=/X (there the reference reads past the sequence)
    (100, [(5, 3), (10, 0), (3, 4)], 13),                       # 5N10M3S
    (140, [(2, 4), (4, 3), (6, 0), (2, 2), (5, 0), (4, 5)], 13),  # 2S4N6M2D5M4H
    (200, [(3, 3), (2, 1), (2, 3), (8, 0), (3, 2)], 10),        # 3N2I2N8M3D: two leading N operations, the tail is counted as deletions
    (260, [(7, 3), (4, 2), (6, 0), (20, 3), (5, 0), (1, 1)], 12),  # 7N4D6M20N5M1I: the first non-N operation is a D
    (330, [(6, 3), (3, 4)], 3),                                 # 6N3S: nothing but N consumes the reference -- plain skips
    (360, [(4, 3), (9, 0), (2, 6)], 9),                         # 4N9M2P
]


def _lead_n_bam(tmp_path, name="lead.bam"):
    import random
    rng = random.Random(5)
    reads, quals = [], []
    for i, (pos, cig, ls) in enumerate(LEAD_N_READS):
        reads.append((0, pos, 30, 0, cig, "".join(rng.choice("ACGT") for _ in range(ls)), f"n{i}"))
        quals.append([rng.choice([5, 30, 40]) for _ in range(ls)])
    # ordinary neighbours so that columns hold more than one read
    for pos in (95, 150, 210, 270, 335, 365):
        reads.append((0, pos, 30, 0, [(30, 0)], "".join(rng.choice("ACGT") for _ in range(30)), f"m{pos}"))
        quals.append([30] * 30)
    order = sorted(range(len(reads)), key=lambda i: reads[i][1])
    return helpers.write_bam(str(tmp_path / name), [("r0", 1000)], [reads[i] for i in order], quals=[quals[i] for i in order])


def test_leading_n_closed_form_equals_the_sweep(tmp_path):
    """Quirk 1 (pileup.d:180-189): the cursor steps over leading N operations without consuming them.  The faithful sweep restates the
    constructor literally; the per-base closed form uses the equivalent CIGAR (closed_form_lead_n) -- they must agree column by column, at
    two -q settings."""
    p = _lead_n_bam(tmp_path)
    for minq in (0, 20):
        rc, out, err = helpers.oracle_cli(["base", "-c", "0", "-q", str(minq), p])
        assert rc == 0, err
        rows = {int(l.split(b"\t")[1]): [int(x) for x in l.split(b"\t")[2:9]] for l in out.splitlines()[1:]}
        counts, _ = helpers.oracle_counts(p, min_bq=minq)
        for pos, r in rows.items():
            col = counts[:, pos]
            assert [int(col.sum()), col[0], col[1], col[2], col[3], col[5], col[6]] == r, (minq, pos)
        # the first read: 5N10M3S at 100 -- bases at 100..109 (not 105..114), then five skipped columns
        assert rows[100][0] >= 1 and rows[112][6] >= 1
    # the closed form of the region statistics does not model such reads (the reference mixes the CIGAR as written with the shifted cursor
    # there; the product refuses them in region / window mode): it says so instead of returning numbers
    import numpy as np
    with pytest.raises(Exception):
        helpers.oracle_segment_stats(p, np.array([90], dtype=np.uint64), np.array([120], dtype=np.uint64), (1,))
