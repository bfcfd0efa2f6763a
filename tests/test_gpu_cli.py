"""Drop-in CLI parity: sambamba-depth-b200 (C++ host over libbdepth.so) must print byte-identical text to
(1) the reference's own golden files (test/test_suite.sh:149-162,176-194) and (2) the oracle CLI."""
import os

import pytest

import helpers
from helpers import GOLDEN

pytestmark = pytest.mark.gpu


def G(f):
    return os.path.join(GOLDEN, f)


def check_same(args):
    rc1, out1, err1 = helpers.run_cli(args)
    rc2, out2, err2 = helpers.oracle_cli(args)
    assert rc1 == rc2, (args, err1, err2)
    assert out1 == out2, (args, out1[:400], out2[:400])
    return out1, err1


def test_golden_issue_193():
    rc, out, err = helpers.run_cli(["depth", "base", G("issue_193.bam")])
    assert rc == 0, err
    assert out == open(G("issue_193_expected_output.txt"), "rb").read()


def test_golden_issue_225():
    for extra in ([], ["-L", "chrM"]):
        rc, out, err = helpers.run_cli(["depth", "base", "-c", "1"] + extra + [G("issue225.bam")])
        assert rc == 0 and out == open(G("issue225.out"), "rb").read(), (extra, err)
        rc, out, err = helpers.run_cli(["depth", "base", "-c", "0"] + extra + [G("issue225.bam")])
        assert rc == 0 and out == open(G("issue225.z.out"), "rb").read(), (extra, err)


def test_region_204_without_m_matches_oracle():
    # the reference's golden file for issue_204 uses -m (tests/test_zz_gpu_mates.py); without -m the oracle is the checker
    out, _ = check_same(["region", G("issue_204.bam"), "-L", "2:166868600-166868813", "-T", "15", "-T", "20", "-T", "25"])
    assert b"41\t20.2196\t89.7196\t61.215\t20.5607" in out


@pytest.fixture(scope="module")
def tiny(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    p = helpers.gen_bam(str(d / "tiny.bam"), "--preset", "tiny", "-t", 4)
    bed = d / "r.bed"
    bed.write_text("ctgA\t100\t900\tgeneA\nctgA\t850\t1200\tgeneB\nctgC\t0\t52000\nctgB\t10\t20\nnope\t1\t5\nctgA\t29000\t31000\tedge\n")
    sbed = d / "s.bed"
    sbed.write_text("ctgA\t100\t900\nctgA\t1000\t1200\nctgB\t10\t20\nctgC\t5\t40000\n")
    return p, str(bed), str(sbed)


def test_base_modes_match_oracle(tiny):
    p, bed, sbed = tiny
    for args in (["base", p], ["base", "-c", "0", p], ["base", "-z", p], ["base", "-q", "25", p], ["base", "-c", "5", "-C", "9", p],
                 ["base", "-a", "-c", "3", p], ["base", "--combined", p], ["base", "-F", "", p], ["base", "-L", "ctgB", p],
                 ["base", "-L", "ctgA:1,000-2000", "-c", "0", p], ["base", "-L", sbed, p], ["base", "-L", bed, "-c", "0", p]):
        check_same(args)


def test_window_modes_match_oracle(tiny):
    p, _, _ = tiny
    for args in (["window", "-w", "1000", p], ["window", "-w", "500", "-T", "5", "-T", "0", "-T", "12", p], ["window", "-w", "777", "-q", "30", p],
                 ["window", "-w", "100000", p], ["window", "-w", "1000", "-c", "8", p], ["window", "-w", "1000", "-a", "-C", "8.2", "--combined", p]):
        check_same(args)


def test_overlapping_windows_match_oracle(tiny):
    # the reference keeps ceil(w/step) ring slots (depth.d:1026-1029); a reused slot collects thresholds before its window
    # starts when step does not divide w (depth.d:215-226), and the first reference's initial slots only count reads
    # that start inside them (is_first_occurrence starts false, depth.d:1031-1032) -- both reproduced in closed form
    p, _, _ = tiny
    for args in (["window", "-w", "1000", "--overlap", "500", p], ["window", "-w", "1000", "--overlap", "300", "-T", "3", "-T", "9", p],
                 ["window", "-w", "1000", "--overlap", "900", "-T", "5", p], ["window", "-w", "777", "--overlap", "100", "-T", "4", "-q", "20", p],
                 ["window", "-w", "640", "--overlap", "639", "-T", "6", "-L", "ctgB", p], ["window", "-w", "1500", "--overlap", "1100", "-T", "2", "-c", "6", "-a", p]):
        check_same(args)


def test_region_modes_match_oracle(tiny):
    p, bed, sbed = tiny
    for args in (["region", "-L", sbed, p], ["region", "-L", sbed, "-T", "3", "-T", "10", p], ["region", "-L", bed, "-T", "8", p],
                 ["region", "-L", "ctgA:5000-6000", "-q", "20", p], ["region", "-L", "ctgB", "-c", "1", p], ["region", "-L", sbed, "-a", "-c", "7.9", p]):
        check_same(args)


def test_errors_and_usage(tiny):
    p, _, _ = tiny
    rc, out, err = helpers.run_cli(["depth"])
    assert rc == 0 and b"Usage: sambamba-depth" in err
    rc, out, err = helpers.run_cli(["region", p])
    assert rc == 1 and b"BED file or a region must be provided" in err
    rc, out, err = helpers.run_cli(["base", "/nonexistent.bam"])
    assert rc == 1 and err.startswith(b"sambamba-depth: ")
    rc, out, err = helpers.run_cli(["window", p])
    assert rc == 1 and b"positive window size must be specified" in err


def test_annotated_rows_of_positions_whose_bases_all_fail_q(tiny, tmp_path):
    """-a with -q and a positive minimum coverage: a position that reads cover but whose every base is below -q still has a
    column; the reference prints it with flag n (depth.d:534-555).  The counters are zero there: a presence bitmap
    (k_presence) tells it from a position without reads."""
    p, _, sbed = tiny
    for args in (["base", "-a", "-q", "38", "-c", "1", p], ["base", "-a", "-q", "41", "-c", "3", "-L", "ctgA:1-3000", p], ["base", "-a", "-q", "39", "-c", "2", "--combined", p],
                 ["base", "-a", "-q", "40", "-c", "1", "-L", sbed, p]):
        out, _ = check_same(args)
        assert b"\t0\t0\t0\t0\t0\t0\t0\t" in out          # rows with COV 0 are there
    import test_emul_mates as tem
    ms = tem.make_pairs_bam(str(tmp_path / "ms.bam"), 3, rg=[("g1", "S1"), ("g2", "S2")])
    check_same(["base", "-a", "-q", "38", "-c", "2", ms])


def test_window_longer_than_the_last_reference_with_reads(tmp_path):
    """The first trailing reference without reads is printed before the window state is reset (depth.d:1070-1076): its first rows carry what
    the ring still holds -- also when the last reference with reads was shorter than one window (no full window there: shift 0).  Found by
    tools/fuzz_emul.py."""
    p = helpers.write_bam(str(tmp_path / "short.bam"), [("r0", 1000), ("r1", 70000), ("r2", 500)],
                          [(0, 328, 52, 145, [(2, 0)], "AC", "q0"), (0, 400, 30, 0, [(50, 0)], "A" * 50, "q1")])
    for args in (["window", "-w", "100000", "--overlap", "98149", "-T", "2", p], ["window", "-w", "2000", p], ["window", "-w", "2000", "--overlap", "1500", "-T", "1", p],
                 ["window", "-w", "1000", p], ["window", "-w", "1001", "-a", p], ["window", "-w", "999", "--overlap", "3", p]):
        out, _ = check_same(args)
    out, _ = check_same(["window", "-w", "100000", "--overlap", "98149", p])
    assert b"r1\t0\t100000\t2\t0.00052" in out, out[:300]


def test_zero_coverage_rows_skip_a_reference_without_reads_between_two_with_reads(tmp_path):
    """--min-coverage=0: the reference writes the empty rows of the references before the first one with reads, behind the last one, and of
    the gaps -- but moving from one reference to a later one it writes only the tail of the former and the head of the latter
    (PerBasePrinter.push, depth.d:578-581): a reference in between that the sweep never sees gets no rows.  "Sees" = a read that
    passes the filter and, with -L, overlaps a region (found by the differential fuzzer, seeds 101-103)."""
    refs = [("r0", 30), ("r1", 20), ("r2", 25), ("r3", 10), ("r4", 12)]
    M = lambda n: [(n, 0)]
    reads = [(1, 5, 30, 0, M(10), "ACGTACGTAC", "a"),
             (2, 2, 0, 0, M(6), "ACGTAC", "f"),            # mapping quality 0: fails the default filter
             (2, 3, 30, 0x400, M(6), "ACGTAC", "g"),       # duplicate: fails it too
             (3, 3, 30, 0, M(4), "ACGT", "b")]
    p = helpers.write_bam(str(tmp_path / "z.bam"), refs, reads)
    for args in (["base", "-z", p], ["base", "-c", "0", "--combined", p], ["base", "-a", "-c", "0", p]):
        out, _ = check_same(args)
        names = [l.split(b"\t")[0] for l in out.splitlines()[1:]]
        assert names.count(b"r0") == 30 and names.count(b"r1") == 20 and names.count(b"r3") == 10 and names.count(b"r4") == 12
        assert names.count(b"r2") == 0              # every read on it fails the filter: skipped between r1 and r3
    out, _ = check_same(["base", "-c", "0", "-F", "", p])     # without the filter r2 has reads: all its rows
    assert [l.split(b"\t")[0] for l in out.splitlines()[1:]].count(b"r2") == 25
    # with -L only reads that overlap a region are seen: r3's read lies outside its region, so r3 is the skipped one now
    bed = tmp_path / "z.bed"
    bed.write_text("r0\t3\t9\nr1\t0\t8\nr2\t0\t5\nr3\t8\t10\nr4\t1\t4\n")
    reads2 = reads + [(4, 2, 30, 0, M(3), "ACG", "c")]
    p2 = helpers.write_bam(str(tmp_path / "z2.bam"), refs, reads2)
    out, err = check_same(["base", "-c", "0", "-F", "", "-L", str(bed), p2])
    names = [l.split(b"\t")[0] for l in out.splitlines()[1:]]
    assert names.count(b"r0") == 6 and names.count(b"r1") == 8 and names.count(b"r2") == 5 and names.count(b"r3") == 0 and names.count(b"r4") == 3
    assert b"(r3)" not in err and b"(r2)" in err      # "Processing reference ..." lists what the sweep saw


def test_cigars_that_begin_with_n(tmp_path):
    """Quirk 1 (pileup.d:180-189): PileupRead steps over leading N operations without consuming them, so the rest of the CIGAR is applied
    that many columns early and the columns left at the end follow the last operation.  k2_lead_n rewrites such CIGARs to the equivalent
    one; every mode must then print what the reference's sweep prints.  A CIGAR that also ENDS in M/=/X sends the reference's cursor past
    the read's sequence: refused with a message."""
    from test_oracle_golden import _lead_n_bam
    p = _lead_n_bam(tmp_path)
    bed = tmp_path / "s.bed"
    bed.write_text("r0\t90\t120\nr0\t120\t215\nr0\t215\t300\nr0\t320\t400\n")
    for args in (["base", p], ["base", "-c", "0", p], ["base", "-q", "20", "-a", p], ["base", "-L", str(bed), p], ["base", "-F", "", "-q", "35", p],
                 ["base", "-m", p], ["base", "-m", "-q", "20", "-L", str(bed), p]):
        check_same(args)
    # region statistics of such a read: the reference takes readCount and meanCoverage from the CIGAR as written (countOverlappingBases,
    # depth.d:671-698) and the percentages from the shifted cursor -- k2_lead_n books the difference (one rank, no -m)
    obed = tmp_path / "o.bed"
    obed.write_text("r0\t90\t120\tx\nr0\t100\t101\ty\nr0\t105\t300\tz\nr0\t265\t272\nr0\t290\t296\nr0\t330\t400\tw\nr0\t0\t1000\n")
    for args in (["region", "-L", str(bed), "-T", "1", "-T", "2", p], ["region", "-L", str(bed), "-q", "20", "-T", "1", p], ["region", "-L", str(obed), "-T", "1", "-T", "3", p],
                 ["region", "-L", str(obed), "-q", "35", "-a", "-c", "0.5", p], ["region", "-L", "r0:263-275", "-F", "", p]):
        check_same(args)
    for args in (["window", "-w", "50", p], ["window", "-w", "40", "-q", "20", "-T", "1", p]):       # windows that do not overlap are regions
        check_same(args)
    # overlapping windows and -m keep per-slot / per-pair books of their own on top of that: refused with a message (reads without a leading N: as always)
    for args in (["window", "-w", "40", "--overlap", "10", p], ["region", "-m", "-L", str(bed), p]):
        rc, out, err = helpers.run_cli(["depth"] + args)
        assert rc != 0 and b"begins with N" in err, (args, err)
    check_same(["region", "-L", "r0:380-500", p])          # (sparse staging or not, the leading-N reads lie in front of this region ...
    rc, out, err = helpers.run_cli(["depth", "window", "-w", "50", "-F", "read_name =~ /^m/", p])      # ... and here they are filtered out)
    assert rc == 0 and out.count(b"\n") == 21, err
    import sambamba_b200 as sb
    for minq in (0, 20):
        want, _ = helpers.oracle_counts(p, min_bq=minq)
        with sb.BDepth(p) as b:
            b.set_min_baseq(minq)
            got = b.run_base()
        assert (got == want).all()
    # the raw record scan and the index builder see the file as it is (the rewrite is for the pileup only)
    with sb.BDepth(p) as b:
        assert b.build_index() == helpers.oracle_build_bai(p)
    # ... ending in a match: refused
    bad = helpers.write_bam(str(tmp_path / "bad.bam"), [("r0", 1000)], [(0, 10, 30, 0, [(20, 0)], "A" * 20, "ok"), (0, 50, 30, 0, [(5, 3), (10, 0)], "ACGTACGTAC", "leadn")])
    rc, out, err = helpers.run_cli(["depth", "base", bad])
    assert rc != 0 and b"begins with N" in err
    rc, out, err = helpers.run_cli(["depth", "base", "-F", "mapping_quality > 40", bad])       # filtered out: not an issue
    assert rc == 0
