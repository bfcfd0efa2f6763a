"""`-m` / --fix-mate-overlaps through the GPU: libbdepth.so's km_hash / km_link / km_fix after K3.

The device functions are the ones tests/test_emul_mates.py runs on the CPU against the oracle's column sweep; what
only hardware can show is the launch plumbing (single-batch mode, parameter block, region arrays).  This file was
written after the round's GPU budget was spent, so its first hardware run is still pending: the tests execute (and
report XPASS when green) but do not gate the suite until they have been seen green once -- then the xfail marker goes.
The file sorts last on purpose: nothing runs after it in the same process."""
import os

import numpy as np
import pytest

import helpers
from helpers import GOLDEN

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def G(f):
    return os.path.join(GOLDEN, f)


def check_same(args):
    rc1, out1, err1 = helpers.run_cli(args)
    rc2, out2, err2 = helpers.oracle_cli(args)
    assert rc1 == rc2, (args, err1, err2)
    assert out1 == out2, (args, out1[:400], out2[:400])
    return out1


@pytest.fixture(scope="module")
def pairs(tmp_path_factory):
    d = tmp_path_factory.mktemp("mates")
    p = helpers.gen_bam(str(d / "pairs.bam"), "--preset", "tiny", "-n", 20000, "--pairs", 6, "-t", 4)
    n_big = 12000 if os.environ.get("BDEPTH_EMULATE") == "1" else 120000       # the CPU emulation of the pipeline is ~1000x slower than the GPU
    big = helpers.gen_bam(str(d / "pairs2.bam"), "-r", "chrA:900000", "-r", "chrB:600000", "-n", n_big, "--pairs", 5, "-s", 9, "-t", 4)
    sbed = d / "s.bed"
    sbed.write_text("ctgA\t100\t900\nctgA\t1000\t1200\nctgA\t1200\t1207\nctgB\t10\t20\nctgC\t5\t40000\n")
    obed = d / "o.bed"
    obed.write_text("ctgA\t100\t900\tgeneA\nctgA\t850\t1200\tgeneB\nctgC\t0\t52000\nctgB\t10\t20\nctgA\t29000\t31000\tedge\n")
    return p, big, str(sbed), str(obed)


def test_golden_issue_204_region_with_m():
    """The reference's only region-mode golden vector (test/test_suite.sh:156-162): byte-identical through the GPU."""
    rc, out, err = helpers.run_cli(["depth", "region", G("issue_204.bam"), "-L", "2:166868600-166868813", "-T", "15", "-T", "20", "-T", "25", "-m"])
    assert rc == 0, err
    assert out == open(G("issue_204_expected_output.txt"), "rb").read()


def test_mate_overlaps_fixture():
    bed = G("mate_overlaps_1_3M_4M.bed")
    check_same(["base", "-m", G("mate_overlaps_1_3M_4M.bam")])
    check_same(["base", "-m", "-c", "0", "-L", bed, G("mate_overlaps_1_3M_4M.bam")])
    check_same(["region", "-m", "-L", bed, "-T", "1", "-T", "2", G("mate_overlaps_1_3M_4M.bam")])


def test_base_counters_match_closed_form(pairs):
    import sambamba_b200 as sb
    for p in pairs[:2]:
        for minq in (0, 25):
            want, npc = helpers.oracle_counts_fix_mates(p, min_bq=minq)
            with sb.BDepth(p) as b:
                b.set_fix_mates(True)
                b.set_min_baseq(minq)
                got = b.run_base()
                st = b.stats()
            assert npc > 1000 and st["mate_pair_columns"] == npc and st["mate_pairs"] > 30
            assert got.shape == want.shape and np.array_equal(got, want)
            # staged input (one K1 launch) takes the same single-batch path
            with sb.BDepth(p) as b:
                b.set_fix_mates(True)
                b.set_min_baseq(minq)
                b.stage()
                assert np.array_equal(b.run_base(), want)


def test_cli_base_and_region_match_oracle(pairs):
    p, _, sbed, obed = pairs
    for args in (["base", "-m", p], ["base", "-m", "-c", "0", "-q", "20", p], ["base", "-m", "--combined", "-a", "-c", "4", p],
                 ["base", "-m", "-L", sbed, "-c", "0", p], ["base", "-m", "-L", "ctgA:1,000-2000", p],
                 ["region", "-m", "-L", sbed, "-T", "3", "-T", "10", p], ["region", "-m", "-L", sbed, "-q", "25", "-T", "0", "-T", "6", p],
                 ["region", "-m", "-L", obed, "-T", "8", p], ["region", "-m", "-L", "ctgA:5000-6000", "-a", "-c", "7.9", p]):
        check_same(args)


def test_multi_sample_pairs(tmp_path):
    import test_emul_mates as tem
    rg = [("g1", "S1"), ("g2", "S2"), ("g3", "S1")]
    p = tem.make_pairs_bam(str(tmp_path / "ms.bam"), 7, rg=rg)
    check_same(["base", "-m", "-c", "0", p])
    check_same(["base", "-m", "--combined", p])
    check_same(["region", "-m", "-L", "c1:100-900", "-T", "2", p])


def test_groups_of_three_and_more(tmp_path, pairs):
    """Supplementary alignments sharing a name: the per-name replay of the reference's state machine, base and region mode."""
    import test_emul_mates as tem
    for seed in range(10, 13):
        p = tem.make_pairs_bam(str(tmp_path / f"tri{seed}.bam"), seed, n_frag=120, triples=0.5)
        check_same(["base", "-m", "-c", "0", "--combined", p])
        check_same(["region", "-m", "-L", "c1:1-4000", "-T", "2", p])
        bed = tmp_path / f"b{seed}.bed"
        bed.write_text("".join(f"c1\t{a}\t{a + w}\n" for a, w in ((50, 7), (57, 200), (300, 1), (400, 90), (490, 600), (1500, 30))) + "c2\t10\t2000\n")
        check_same(["region", "-m", "-L", str(bed), "-T", "1", "-T", "4", p])


def test_window_mode(pairs, tmp_path):
    import test_emul_mates as tem
    p = pairs[0]
    for args in (["window", "-w", "1000", "-m", p], ["window", "-w", "777", "-m", "-T", "3", "-T", "9", "-q", "20", p], ["window", "-w", "64", "-m", "--combined", "-a", "-c", "5", p]):
        check_same(args)
    q = tem.make_pairs_bam(str(tmp_path / "w.bam"), 41, n_frag=500, triples=0.2)
    check_same(["window", "-w", "100", "-m", "-T", "2", q])
    # overlapping windows: ring slots updated before their window begins, first-occurrence quirk of reference 0
    for w, o in ((100, 50), (100, 30), (90, 80), (64, 63)):
        check_same(["window", "-w", str(w), "--overlap", str(o), "-m", "-T", "2", "-T", "5", q])
    check_same(["window", "-w", "1000", "--overlap", "300", "-m", "-T", "3", p])


def test_several_batches(pairs, tmp_path):
    """-m across batch boundaries: 1 MB batches cut the pair stream many times; the counters, the pair statistics and the
    region / window tables must equal the one-batch run and the oracle."""
    import sambamba_b200 as sb
    import test_emul_mates as tem
    p = pairs[0]
    want, npc = helpers.oracle_counts_fix_mates(p)
    regs = [(0, 100, 900), (0, 1000, 1200), (0, 1200, 1207), (1, 10, 20), (2, 5, 40000)]
    with sb.BDepth(p) as b:
        b.set_fix_mates(True)
        one = b.run_base()
        st1 = b.stats()
        r1 = b.run_regions(regs, [3, 10])
        w1 = b.run_windows(700, 300, [2])
        b.set_tuning(1 << 20, 0)
        many = b.run_base()
        st = b.stats()
        assert st["n_batches"] >= 4 and st1["n_batches"] == 1
        assert np.array_equal(one, want) and np.array_equal(many, want)
        assert st["mate_pair_columns"] == npc == st1["mate_pair_columns"] and st["mate_pairs"] == st1["mate_pairs"] and st["n_records"] == st1["n_records"]
        assert b.run_regions(regs, [3, 10]) == r1
        assert b.run_windows(700, 300, [2]) == w1
        for batch in (1 << 16, 200000, 3 << 16):      # down to one BGZF block per batch: a boundary every 64 KB
            b.set_tuning(batch, 0)
            assert np.array_equal(b.run_base(), want), batch
            stb = b.stats()
            assert stb["mate_pair_columns"] == npc and stb["n_records"] == st1["n_records"]
        b.set_tuning(1 << 16, 0)
        assert b.run_regions(regs, [3, 10]) == r1
    # names with three and more reads, dense, in small batches
    q = tem.make_pairs_bam(str(tmp_path / "tri.bam"), 91, n_frag=4000, refs=(("c1", 6000), ("c2", 2500)), triples=0.3)
    want, npc = helpers.oracle_counts_fix_mates(q)
    with sb.BDepth(q) as b:
        b.set_fix_mates(True)
        for batch in (1 << 20, 1 << 16):
            b.set_tuning(batch, 0)
            got = b.run_base()
            st = b.stats()
            assert st["n_batches"] >= 2 and st["mate_groups"] > 50
            assert np.array_equal(got, want), batch


def test_long_chains_of_one_name(tmp_path):
    """A chain of overlapping reads of one name may be any length (the walk holds only the members of the current column); nine reads of a
    name over one position are refused with a message.  Several batches: a chain cut by batch boundaries is fixed in the batch that closes it."""
    import sambamba_b200 as sb
    import test_emul_mates as tem
    for seed in (300, 301):
        p = tem.make_pairs_bam(str(tmp_path / f"ch{seed}.bam"), seed, n_frag=60, triples=0.3, chains=0.3)
        check_same(["base", "-m", "-c", "0", "--combined", p])
        check_same(["region", "-m", "-L", "c1:1-4000", "-T", "2", p])
        check_same(["window", "-w", "100", "-m", "-T", "2", p])
    q = tem.make_pairs_bam(str(tmp_path / "chb.bam"), 77, n_frag=1500, refs=(("c1", 60000), ("c2", 2500)), triples=0.2, chains=0.2)
    want, npc = helpers.oracle_counts_fix_mates(q)
    with sb.BDepth(q) as b:
        b.set_fix_mates(True)
        assert np.array_equal(b.run_base(), want)
        b.set_tuning(1 << 16, 0)
        got = b.run_base()
        assert b.stats()["n_batches"] >= 2 and np.array_equal(got, want)
    p9 = tem.make_pairs_bam(str(tmp_path / "p9.bam"), 5, n_frag=20, pile=9)
    rc, out, err = helpers.run_cli(["depth", "base", "-m", p9])
    assert rc == 1 and b"more than 8 reads of one name cover one position" in err, err
    p8 = tem.make_pairs_bam(str(tmp_path / "p8.bam"), 5, n_frag=20, pile=8)
    check_same(["base", "-m", "-c", "0", p8])
