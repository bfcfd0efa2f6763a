"""BAI builder on the GPU (`sambamba index`, BioD/bio/std/hts/bam/bai/indexing.d:56-366): bdepth_build_index against
the oracle's restatement of IndexBuilder (byte for byte: both write a reference's bins in ascending order) and against the .bai files the
reference ships next to its test BAMs (as parsed structures: the reference writes bins in the order of a D associative array)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(helpers.ROOT, "tests", "golden")
FIXTURES = ["issue225.bam", "issue_193.bam", "issue_204.bam", "mate_overlaps_1_3M_4M.bam"]


def _build(path, tuning=None):
    import sambamba_b200 as sb
    with sb.BDepth(path) as b:
        if tuning:
            b.set_tuning(*tuning)
        return b.build_index()


@pytest.mark.parametrize("name", FIXTURES)
def test_reference_fixtures(name):
    p = os.path.join(GOLDEN, name)
    got = _build(p)
    assert got == helpers.oracle_build_bai(p)
    assert helpers.parse_bai(got) == helpers.parse_bai(open(p + ".bai", "rb").read()), "the index sambamba itself wrote for this file"


@pytest.mark.parametrize("tuning", [None, (1 << 20, 3), (1 << 18, 1)])
def test_synthetic_multi_reference_across_batches(tmp_path, tuning):
    """Several references (one tiny, one without reads in front), records straddling BGZF members, batches and sub-batches."""
    big = os.environ.get("BDEPTH_EMULATE") != "1"
    p = helpers.gen_bam(str(tmp_path / "s.bam"), "-r", "chr0:5000", "-r", "chrA:%d" % (2000000 if big else 200000), "-r", "chrB:700", "-r", "chrC:%d" % (1500000 if big else 150000),
                        "-n", 300000 if big else 30000, "-s", 12, "-t", 8)
    assert _build(p, tuning) == helpers.oracle_build_bai(p)


def _edge_reads():
    M, I, D, N, S = 0, 1, 2, 3, 4
    seq = "ACGTACGTAC" * 5
    r = []
    # ref 1 (ref 0 stays empty): reads in one leaf bin, one crossing a 16 kbp border (parent bin), a read with a long skip over many windows
    r.append((1, -1, 0, 0x4, [], seq, "placed_unmapped_without_position"))            # counted in the metadata only (indexing.d:298-299)
    r.append((1, 100, 60, 0, [(50, M)], seq, "a1"))
    r.append((1, 120, 60, 0x4, [(50, M)], seq, "a2_unmapped_with_position"))          # basesCovered = 0 (read.d:255-259): one window
    r.append((1, 16000, 60, 0, [(20, M), (400, D), (30, M)], seq, "a3_cross"))
    r.append((1, 16380, 60, 0, [(50, M)], seq, "a4_cross"))
    r.append((1, 16390, 60, 0, [(50, M)], seq, "a5"))
    r.append((1, 20000, 60, 0, [(10, M), (70000, N), (40, M)], seq, "a6_skip"))
    r.append((1, 20000, 0, 0x400, [(5, S), (45, M)], seq, "a7"))
    r.append((1, 49152, 60, 0, [], seq, "a8_no_cigar_at_window_start"))                # mapped flag, nothing covered: end window = start window - 1
    r.append((1, 90000, 60, 0, [(50, M)], seq, "a9"))
    # ref 3 (ref 2 empty in the middle)
    r.append((3, 0, 60, 0, [(50, M)], seq, "c1"))
    r.append((3, 10, 60, 0, [(25, M), (3, I), (22, M)], seq, "c2"))
    r.append((3, 40000, 60, 0, [(50, M)], seq, "c3"))
    # reads without reference at the end (n_no_coor)
    for i in range(5):
        r.append((-1, -1, 0, 0x4, [], seq, "u%d" % i))
    return [("r0", 1000), ("r1", 120000), ("r2", 5000), ("r3", 60000), ("r4", 300)], r


@pytest.mark.parametrize("block", [0xFF00, 211])
def test_edge_cases(tmp_path, block):
    """Unplaced and position-less reads, empty references before / between / after, zero-length alignments, a skip over several linear
    windows, and (block = 211) BGZF members smaller than a record: chunk ends at member borders, merging of a bin's chunks by member."""
    refs, reads = _edge_reads()
    p = helpers.write_bam(str(tmp_path / "e.bam"), refs, reads, block=block, bins="auto", index=False)
    got = _build(p)
    assert got == helpers.oracle_build_bai(p)
    parsed, no_coor = helpers.parse_bai(got)
    assert no_coor == 5 and [len(b) for b, _ in parsed] == [0, len(parsed[1][0]), 0, len(parsed[3][0]), 0]
    assert parsed[1][0][37450][1] == (8, 2)          # mapped, unmapped of r1: the position-less read counts here


def test_bins_are_taken_from_the_records(tmp_path):
    """The reference indexes by the bin field the record carries (indexing.d:108, :222), right or wrong."""
    refs, reads = _edge_reads()
    bins = [4681 + (i % 3) for i in range(len(reads))]
    p = helpers.write_bam(str(tmp_path / "b.bam"), refs, reads, bins=bins, index=False)
    assert _build(p) == helpers.oracle_build_bai(p)


def test_unsorted_input_is_refused(tmp_path):
    import sambamba_b200 as sb
    seq = "ACGT" * 10
    reads = [(0, 500, 60, 0, [(40, 0)], seq, "x1"), (0, 100, 60, 0, [(40, 0)], seq, "x2")]
    p = helpers.write_bam(str(tmp_path / "u.bam"), [("r0", 10000)], reads, bins="auto", index=False)
    with pytest.raises(RuntimeError):
        helpers.oracle_build_bai(p)
    with sb.BDepth(p) as b:
        with pytest.raises(sb.BDepthError, match="not coordinate-sorted"):
            b.build_index()


def test_unindexed_input_runs_after_the_index_was_built(tmp_path):
    """A file that came without .bai: the handle adopts the index it built -- depth base equals the oracle, a region query stages sparsely."""
    import sambamba_b200 as sb
    big = os.environ.get("BDEPTH_EMULATE") != "1"
    src = helpers.gen_bam(str(tmp_path / "src.bam"), "-r", "chrA:%d" % (1000000 if big else 120000), "-r", "chrB:%d" % (400000 if big else 60000), "-n", 120000 if big else 15000, "-s", 5, "-t", 8)
    p = str(tmp_path / "noidx.bam")
    shutil.copy(src, p)
    want, _ = helpers.oracle_counts(src)
    with sb.BDepth(p) as b:
        assert not b.has_index
        bai = b.build_index()
        assert b.has_index and bai == helpers.oracle_build_bai(src)
        got = b.run_base()
        assert np.array_equal(got, want)
        regs = [(0, 5000, 5600), (1, 100, 900)]
        rows = b.run_regions(regs, [1])
        st = b.stats()
        assert st["file_bytes"] < os.path.getsize(p) // 2, "the region query only staged the chunks the built index names"
        for (ref, s, e), row in zip(regs, rows):
            lin0 = 0 if ref == 0 else (1000000 if big else 120000)
            assert row[4] == int(want[:5, lin0 + s:lin0 + e].sum())


def test_cli_index_subcommand(tmp_path):
    src = os.path.join(GOLDEN, "issue_204.bam")
    p = str(tmp_path / "i.bam")
    shutil.copy(src, p)
    r = subprocess.run([helpers.CLI, "index", p], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(p + ".bai", "rb").read() == helpers.oracle_build_bai(src)
    out = str(tmp_path / "other.bai")
    r = subprocess.run([helpers.CLI, "index", "-t", "4", p, out], capture_output=True, text=True)
    assert r.returncode == 0 and open(out, "rb").read() == open(p + ".bai", "rb").read()
    os.remove(p + ".bai")
    # depth refuses un-indexed input as the reference does (depth.d:1166) unless asked to index it first
    r = subprocess.run([helpers.CLI, "depth", "base", p], capture_output=True, text=True)
    assert r.returncode == 1 and "must be indexed" in r.stderr
    r2 = subprocess.run([helpers.CLI, "depth", "base", "--build-index", p], capture_output=True, text=True)
    want = subprocess.run([helpers.CLI, "depth", "base", src], capture_output=True, text=True)
    assert r2.returncode == 0 and r2.stdout == want.stdout and len(want.stdout) > 1000


def _random_case(seed):
    """A small random BAM: references with and without reads, position-less reads in front of a reference's reads, unmapped reads with a
    position, empty / skipping / clipped CIGARs, right or arbitrary bin fields, unplaced reads at the end, BGZF members from 150 B to 64 KB."""
    import random
    rnd = random.Random(seed)
    M, I, D, N, S = 0, 1, 2, 3, 4
    refs = [("r%d" % i, rnd.choice([300, 5000, 40000, 200000])) for i in range(rnd.randint(1, 6))]
    reads = []
    for r, (_, ln) in enumerate(refs):
        if rnd.random() < 0.25:
            continue
        pos = sorted(rnd.randrange(0, max(1, ln - 100)) for _ in range(rnd.randint(1, 400)))
        if rnd.random() < 0.3:
            pos = [-1] * rnd.randint(1, 3) + pos
        for p in pos:
            k = rnd.random()
            cig = ([(40, M)] if k < 0.6 else [(10, M), (rnd.randint(1, 30000), N), (30, M)] if k < 0.7 else [(5, S), (20, M), (3, I), (12, M)] if k < 0.8
                   else [(20, M), (rnd.randint(1, 50), D), (20, M)] if k < 0.9 else [] if k < 0.95 else [(rnd.randint(1, 100), M)])
            if p >= 0 and p + sum(l for l, o in cig if o in (0, 2, 3)) > ln:
                cig = [(min(40, max(1, ln - p)), M)]
            reads.append((r, p, rnd.randint(0, 60), 4 if rnd.random() < 0.1 else 0, cig, "ACGT" * 10, "q%d" % len(reads)))
    reads += [(-1, -1, 0, 4, [], "ACGT" * 10, "u%d" % i) for i in range(rnd.randint(0, 5))]
    bins = "auto" if rnd.random() < 0.7 else [rnd.choice([4681, 4682, 585, 73, 0, 4690]) for _ in reads]
    return refs, reads, bins, rnd.choice([0xFF00, 4096, 777, 150]), rnd.choice([None, (1 << 16, 1), (1 << 16, 2), (1 << 17, 3)])


def test_random_files(tmp_path):
    import sambamba_b200 as sb
    for seed in range(40):
        refs, reads, bins, block, tuning = _random_case(seed)
        if not reads:
            continue
        p = helpers.write_bam(str(tmp_path / ("f%d.bam" % seed)), refs, reads, block=block, bins=bins, index=False)
        try:
            want = helpers.oracle_build_bai(p)
        except RuntimeError:
            want = None                      # (a position-less read behind reads of its reference: "not coordinate-sorted", indexing.d:259-271)
        try:
            got = _build(p, tuning)
        except sb.BDepthError:
            got = None
        assert got == want, (seed, block, tuning)
