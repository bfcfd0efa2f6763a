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
