"""Several BAM files in one run (MultiBamReader, BioD/bio/std/hts/bam/multireader.d:218-268; depth.d:1162-1181): the CLI on
[a.bam, b.bam] must print what the oracle prints for ONE file holding the coordinate-sorted union of their records under a
header with both files' read groups.  The reference has no test of its own for this (unpinned)."""
import os
import struct

import numpy as np
import pytest

import helpers

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

REFS = [("chrA", 120000), ("chrB", 600), ("chrC", 80000)]


def _records(path):
    u = helpers.oracle_inflate(path)
    first, refs = helpers.header_first_record_offset(u)
    b = u.tobytes()
    out, o = [], first
    while o + 4 <= len(b):
        bs = struct.unpack_from("<i", b, o)[0]
        if o + 4 + bs > len(b):
            break
        ref, pos = struct.unpack_from("<ii", b, o + 4)
        out.append(((ref if ref >= 0 else 1 << 30), pos, b[o:o + 4 + bs]))
        o += 4 + bs
    return b[:first], refs, out


def merge_bams(paths, dst, rg_lines=None):
    """The nWayUnion of the files' records (stable: ties keep file order) behind the first file's header, whose @RG lines are
    replaced by rg_lines when given."""
    head, refs, recs = None, None, []
    for fi, p in enumerate(paths):
        h, r, rr = _records(p)
        if head is None:
            head, refs = h, r
        recs += [(a, b2, fi, i, raw) for i, (a, b2, raw) in enumerate(rr)]
    recs.sort(key=lambda t: (t[0], t[1], t[2], t[3]))
    if rg_lines is not None:
        l_text = struct.unpack_from("<i", head, 4)[0]
        text = head[8:8 + l_text].decode()
        text = "".join(l + "\n" for l in text.split("\n") if l and not l.startswith("@RG")) + "".join(l + "\n" for l in rg_lines)
        head = b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + head[8 + l_text:]
    return helpers.write_bgzf(dst, head + b"".join(t[4] for t in recs), len(refs))


@pytest.fixture(scope="module")
def pair(tmp_path_factory):
    d = tmp_path_factory.mktemp("mb")
    ra = sum((["-r", f"{n}:{l}"] for n, l in REFS), [])
    a = helpers.gen_bam(str(d / "a.bam"), *ra, "-n", 9000, "-s", 3, "-t", 2)
    b = helpers.gen_bam(str(d / "b.bam"), *ra, "-n", 5000, "-s", 4, "-t", 2)
    m = merge_bams([a, b], str(d / "merged.bam"))
    return a, b, m


def test_base_window_region_equal_the_merged_file(pair, tmp_path):
    a, b, m = pair
    bed = str(tmp_path / "r.bed")
    open(bed, "w").write("chrA\t100\t900\nchrA\t5000\t5400\nchrC\t70000\t71000\n")
    for args in (["base"], ["base", "-c", "0", "-q", "20"], ["window", "-w", "500", "-T", "3", "-T", "12"], ["window", "-w", "300", "--overlap", "100"],
                 ["region", "-L", bed, "-T", "5"], ["base", "-L", "chrC:1000-3000"], ["base", "--combined"]):
        rc1, out1, err1 = helpers.run_cli(args + [a, b])
        rc2, out2, _ = helpers.oracle_cli(args + [m])
        assert rc1 == 0 and rc2 == 0, (args, err1)
        assert out1 == out2, args


def test_samples_of_several_files(tmp_path):
    refs = [("c1", 5000)]
    mk = lambda n, pos0, rgid: [(0, pos0 + 7 * i, 60, 0, [(50, 0)], "ACGTTGCA" * 6 + "AC", f"{rgid}r{i}") for i in range(n)]
    tag = lambda rgid, n: [b"RGZ" + rgid.encode() + b"\0"] * n
    a = helpers.write_bam(str(tmp_path / "a.bam"), refs, mk(40, 100, "x"), rg=[("x", "SA")], tags=tag("x", 40))
    b = helpers.write_bam(str(tmp_path / "b.bam"), refs, mk(30, 150, "y"), rg=[("y", "SB"), ("z", "SA")], tags=tag("y", 30))
    m = merge_bams([a, b], str(tmp_path / "m.bam"), rg_lines=["@RG\tID:x\tSM:SA", "@RG\tID:y\tSM:SB", "@RG\tID:z\tSM:SA"])      # samples in order of first appearance: SA, SB
    for args in (["base"], ["base", "--combined"], ["window", "-w", "100"], ["base", "-c", "0"]):
        rc1, out1, err1 = helpers.run_cli(args + [a, b])
        rc2, out2, _ = helpers.oracle_cli(args + [m])
        assert rc1 == 0 and rc2 == 0, (args, err1)
        assert out1 == out2, args


def test_counters_through_the_abi_and_refusals(pair, tmp_path):
    import sambamba_b200 as sb
    a, b, m = pair
    want, ost = helpers.oracle_counts(m)
    with sb.BDepth(a) as h:
        h.add_input(b)
        got = h.run_base()
        st = h.stats()
        assert np.array_equal(got, want) and st["n_records"] == ost.n_records and st["n_records_pass"] == ost.n_pass
        h.set_fix_mates(True)
        with pytest.raises(sb.BDepthError, match="several BAM files"):
            h.run_base()
    other = helpers.gen_bam(str(tmp_path / "o.bam"), "-r", "chrA:120000", "-r", "chrZ:600", "-n", 100, "-s", 1, "-t", 1)
    with sb.BDepth(a) as h:
        with pytest.raises(sb.BDepthError, match="reference sequences differ"):
            h.add_input(other)
    rc, out, err = helpers.run_cli(["base", a, str(tmp_path / "missing.bam")])
    assert rc == 1 and b"Cannot open file" in err
