"""Edge cases the reference's fixtures do not cover (no N/P/=/X, no stored blocks, no odd framing):
hand-made BAMs through the GPU engine vs the oracle (counters and CLI text)."""
import os
import random

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def rnd_seq(r, n):
    return "".join(r.choice("ACGTN" if r.random() < 0.02 else "ACGT") for _ in range(n))


def compare(path, minq=0, **flt):
    import sambamba_b200 as sb
    want, ost = helpers.oracle_counts(path, min_bq=minq, **flt)
    with sb.BDepth(path) as b:
        if flt:
            b.set_filter(flt.get("mapq_gt", 0), flt.get("flag_reject", 0x600))
        b.set_min_baseq(minq)
        got = b.run_base()
        st = b.stats()
    assert got.shape == want.shape
    bad = np.argwhere(got != want)
    assert bad.size == 0, (path, bad[:4])
    assert st["n_records"] == ost.n_records and st["n_records_pass"] == ost.n_pass
    return st


def cli_same(args):
    rc1, out1, err1 = helpers.run_cli(args)
    rc2, out2, err2 = helpers.oracle_cli(args)
    assert rc1 == rc2 and out1 == out2, (args, err1[:300], out1[:300], out2[:300])


def test_header_only_and_all_filtered(tmp_path):
    p = helpers.write_bam(str(tmp_path / "empty.bam"), [("c1", 500), ("c2", 300)], [])
    compare(p)
    for args in (["base", p], ["base", "-c", "0", p], ["window", "-w", "100", p], ["region", "-L", "c2:10-200", p]):
        cli_same(args)
    r = random.Random(1)
    reads = [(0, 10 * i, 0, 0, [(50, 0)], rnd_seq(r, 50), f"q{i}") for i in range(20)]          # MAPQ 0
    reads += [(0, 300 + i, 60, 0x400, [(50, 0)], rnd_seq(r, 50), f"d{i}") for i in range(5)]     # duplicates
    reads += [(-1, -1, 0, 4, [], rnd_seq(r, 30), f"u{i}") for i in range(5)]                      # unmapped tail
    p = helpers.write_bam(str(tmp_path / "filtered.bam"), [("c1", 500), ("c2", 300)], reads)
    st = compare(p)
    assert st["n_records"] == 30 and st["n_records_pass"] == 0
    compare(p, mapq_gt=-1, flag_reject=0)
    cli_same(["base", "-F", "", p])
    cli_same(["window", "-w", "64", "-F", "", p])


def test_every_cigar_op_and_clipping_at_reference_end(tmp_path):
    r = random.Random(2)
    L = 4000
    reads = []
    specs = [
        [(30, 0)], [(10, 4), (40, 0)], [(5, 5), (20, 0), (3, 1), (20, 0), (4, 5)], [(20, 0), (7, 2), (30, 0)], [(25, 0), (300, 3), (25, 0)],
        [(10, 7), (5, 8), (10, 7)], [(10, 0), (2, 6), (10, 0)], [(15, 0), (1500, 3), (15, 0), (2, 2), (10, 0)], [(50, 0), (4, 4)],
        [(12, 0), (1, 1), (12, 0), (1, 2), (12, 0), (1, 1), (12, 0)],
    ]
    pos = 5
    for i, cg in enumerate(specs * 6):
        qlen = sum(l for l, op in cg if op in (0, 1, 4, 7, 8))
        reads.append((0, pos, 30 + i % 30, 0x10 if i % 3 else 0, cg, rnd_seq(r, qlen), f"r{i}"))
        pos += r.randint(0, 60)
    # reads hanging over the end of the reference are clipped by both implementations
    reads.append((0, L - 20, 60, 0, [(50, 0)], rnd_seq(r, 50), "over1"))
    reads.append((0, L - 5, 60, 0, [(10, 0), (30, 2), (10, 0)], rnd_seq(r, 20), "over2"))
    reads.sort(key=lambda x: x[1])
    quals = [[r.randint(2, 41) for _ in x[5]] for x in reads]
    p = helpers.write_bam(str(tmp_path / "ops.bam"), [("c1", L), ("c2", 100)], reads, quals=quals)
    st = compare(p)
    assert st["long_reads"] >= 6           # the 1500N reads take the scatter path
    compare(p, minq=20)
    # text parity needs valid input: the reference's sweep prints columns PAST the reference end for overhanging
    # reads (no bounds check, release build), which the engine clips
    keep = [i for i, x in enumerate(reads) if not x[6].startswith("over")]
    p2 = helpers.write_bam(str(tmp_path / "ops_valid.bam"), [("c1", L), ("c2", 100)], [reads[i] for i in keep], quals=[quals[i] for i in keep])
    for args in (["base", p2], ["base", "-q", "25", "-c", "0", p2], ["window", "-w", "250", "-T", "2", p2], ["region", "-L", "c1:100-900", "-q", "10", p2]):
        cli_same(args)


def test_long_reads_spanning_many_bgzf_blocks(tmp_path):
    r = random.Random(3)
    reads = []
    pos = 100
    for i in range(12):
        ops = []
        n = r.randint(3000, 9000)
        for k in range(n):
            ops.append((r.randint(5, 40), 0))
            ops.append((r.randint(1, 3), r.choice([1, 2])))
        ops.append((20, 0))
        qlen = sum(l for l, op in ops if op in (0, 1, 4, 7, 8))
        reads.append((0, pos, 60, 0, ops, rnd_seq(r, qlen), f"long{i}"))       # ~150-400 kb records, > 65535 would need CG; stay below
        pos += r.randint(1000, 50000)
    reads = [x for x in reads if len(x[4]) < 65535]
    p = helpers.write_bam(str(tmp_path / "long.bam"), [("chrL", 2_000_000)], reads, block=0x4000)   # small blocks: every record spans many
    st = compare(p)
    assert st["long_reads"] == len(reads)
    import sambamba_b200 as sb
    want, _ = helpers.oracle_counts(p)
    with sb.BDepth(p) as b:
        b.set_tuning(batch_bytes=1 << 20, chunk_blocks=2)          # force the carry path: records straddle batches and sub-batches
        got = b.run_base()
        assert b.stats()["n_batches"] >= 2
    assert np.array_equal(got, want)


def test_odd_block_sizes_and_levels(tmp_path):
    r = random.Random(4)
    reads = [(i % 2, 50 + 3 * (i // 2), 60, 0, [(40, 0)], rnd_seq(r, 40), f"r{i}") for i in range(3000)]
    reads.sort(key=lambda x: (x[0], x[1]))
    for block, level in ((1, 6), (37, 1), (300, 9), (65280, 0), (4096, 6)):
        if block == 1:
            reads_ = reads[:40]
        else:
            reads_ = reads
        p = helpers.write_bam(str(tmp_path / f"b{block}.bam"), [("a", 5000), ("b", 5000)], reads_, block=max(block, 1), level=level)
        compare(p)


def test_multi_sample_header(tmp_path):
    """@RG -> sample table (depth.d:1170-1181), per-read RG lookup (depth.d:240-250), per-sample rows incl. the
    `return`-not-`continue` quirk of writeColumn (depth.d:540-541), --combined."""
    r = random.Random(7)
    rg = [("rgA", "S1"), ("rgB", "S1"), ("rgC", "S2"), ("rgD", "S3")]
    reads, tags = [], []
    pos = 3
    for i in range(900):
        cg = [(40, 0)] if i % 7 else [(15, 0), (3, 2), (25, 0)]
        reads.append((i % 2 if i > 450 else 0, pos % 2500, 60, 0, cg, rnd_seq(r, 40), f"m{i}"))
        k = i % 9
        tags.append(b"NMC\x01" + (b"" if k == 8 else b"RGZ" + rg[k % 4][0].encode() + b"\0") + b"ASC\x20")
        pos += r.randint(0, 9)
    order = sorted(range(len(reads)), key=lambda i: (reads[i][0], reads[i][1]))
    reads = [reads[i] for i in order]
    tags = [tags[i] for i in order]
    p = helpers.write_bam(str(tmp_path / "ms.bam"), [("c1", 3000), ("c2", 3000), ("c3", 400)], reads, rg=rg, tags=tags)
    import sambamba_b200 as sb
    with sb.BDepth(p) as b:
        assert b.samples == ["S1", "S2", "S3"]
        per = b.run_base()
        assert per.shape[0] == 3
        b.set_combined(True)
        comb = b.run_base()
    want, _ = helpers.oracle_counts(p)
    assert np.array_equal(comb, want)
    assert np.array_equal(per.sum(axis=0), want)
    bed = tmp_path / "ms.bed"
    bed.write_text("c1\t10\t500\nc2\t100\t900\tx\nc1\t400\t1200\n")
    for args in (["base", p], ["base", "-c", "4", p], ["base", "-c", "0", "-L", "c1:100-300", p], ["base", "-a", "-c", "6", p], ["base", "--combined", p],
                 ["window", "-w", "500", "-T", "3", p], ["window", "-w", "500", "--combined", "-c", "2", p], ["region", "-L", str(bed), "-T", "2", "-T", "9", p],
                 ["region", "-L", "c2", "-a", "-c", "3", p]):
        cli_same(args)
    # a read group that is not in the header is an error in both implementations
    tags2 = list(tags)
    tags2[5] = b"RGZnope\0"
    p2 = helpers.write_bam(str(tmp_path / "bad.bam"), [("c1", 3000), ("c2", 3000), ("c3", 400)], reads, rg=rg, tags=tags2)
    rc1, _, err1 = helpers.run_cli(["base", p2])
    rc2, _, err2 = helpers.oracle_cli(["base", p2])
    assert rc1 == 1 and rc2 == 1 and b"read group" in err1 and b"not present in the header" in err2


def test_overlapping_windows_first_reference_without_reads(tmp_path):
    # reads only on the 2nd and 3rd reference: the empty first reference is printed through printEmptyWindows
    # (depth.d:1037-1042), which also resets the slots, so the first-occurrence quirk of reference 0 does not apply
    r = random.Random(11)
    reads = []
    for ref in (1, 2):
        for i in range(600):
            n = r.randint(30, 120)
            reads.append((ref, r.randint(0, 5000 - 130), 60, 0, [(n, 0)], rnd_seq(r, n), f"r{ref}_{i}"))
    reads.sort(key=lambda x: (x[0], x[1]))
    p = helpers.write_bam(str(tmp_path / "w.bam"), [("e0", 3000), ("c1", 5000), ("c2", 5000), ("e3", 2500)], reads)
    for args in (["window", "-w", "400", "--overlap", "100", "-T", "3", p], ["window", "-w", "400", "--overlap", "200", p],
                 ["window", "-w", "250", "--overlap", "249", "-T", "10", "-L", "c2", p]):
        cli_same(args)
    # and with reads on reference 0 (quirk applies there only)
    reads0 = [(0, r.randint(0, 2800), 60, 0, [(100, 0)], rnd_seq(r, 100), f"z{i}") for i in range(300)]
    reads0.sort(key=lambda x: x[1])
    p2 = helpers.write_bam(str(tmp_path / "w0.bam"), [("e0", 3000), ("c1", 5000), ("c2", 5000), ("e3", 2500)], reads0 + reads)
    for args in (["window", "-w", "400", "--overlap", "100", "-T", "3", p2], ["window", "-w", "330", "--overlap", "300", "-T", "2", "-q", "15", p2]):
        cli_same(args)


def test_corrupt_and_truncated_files_are_refused(tmp_path):
    """Malformed input must end in an error message, never in a hang, a fault or quietly shortened output: a truncated last
    record (readExact throws in the reference, readrange.d:169), a damaged DEFLATE stream (zlib error, block.d:127-216), a
    record whose fields overrun its block_size (the release build of the reference slices past it unchecked; this engine
    refuses), a cut file.  Each case runs in its own process (the CLI), so a device fault could not hide behind a later test."""
    import struct
    import numpy as np
    src = helpers.gen_bam(str(tmp_path / "src.bam"), "-r", "chrA:100000", "-r", "chrC:50000", "-n", 8000, "-s", 3, "-t", 2)
    u = helpers.oracle_inflate(src)
    first, _ = helpers.header_first_record_offset(u)
    recs = helpers.parse_records(u, first)
    o = recs[len(recs) // 2][0]
    body = bytearray(u.tobytes())

    def body_with(fn):
        v = bytearray(body)
        fn(v)
        return helpers.write_bgzf(str(tmp_path / "c.bam"), bytes(v), 2)

    cases = [
        ("block_size 0", lambda v: struct.pack_into("<i", v, o, 0)),
        ("block_size beyond the file", lambda v: struct.pack_into("<i", v, o, 0x7FFFFFF0)),
        ("n_cigar_op overruns the record", lambda v: struct.pack_into("<H", v, o + 16, 65535)),
        ("l_seq overruns the record", lambda v: struct.pack_into("<i", v, o + 20, 0x7FFFFFF0)),
        ("negative l_seq", lambda v: struct.pack_into("<i", v, o + 20, -3)),
    ]
    for what, fn in cases:
        for args in (["base"], ["base", "-q", "20"], ["window", "-w", "1000"]):
            rc, out, err = helpers.run_cli(args + [body_with(fn)])
            assert rc != 0 and err.strip(), (what, args, rc, err[:200])
    p = helpers.write_bgzf(str(tmp_path / "cut.bam"), bytes(body[:len(body) - 17]), 2)      # the stream ends inside the last record
    rc, out, err = helpers.run_cli(["base", p])
    assert rc != 0 and b"not enough data" in err, err[:200]
    rc2, _, err2 = helpers.oracle_cli(["base", p])
    assert rc2 != 0 and b"not enough data" in err2, err2[:200]
    p = helpers.write_bgzf(str(tmp_path / "cut2.bam"), bytes(body[:recs[-1][0] + 2]), 2)     # two stray bytes after the last record: a quiet end (readrange.d:139-149)
    rc, out, err = helpers.run_cli(["base", p])
    assert rc == 0, err[:200]
    raw = open(src, "rb").read()
    bad = tmp_path / "deflate.bam"
    flipped = 0
    for pos in range(len(raw) // 2, len(raw) // 2 + 4000, 97):        # somewhere in there a flip breaks a Huffman stream
        data = bytearray(raw)
        data[pos] ^= 0x55
        bad.write_bytes(bytes(data))
        rc, out, err = helpers.run_cli(["base", str(bad)])
        assert rc == 0 or err.strip(), (pos, rc)
        flipped += rc != 0
    assert flipped, "no flip was detected"
    for cut in (len(raw) // 2, 100, 10):
        bad.write_bytes(raw[:cut])
        rc, out, err = helpers.run_cli(["base", str(bad)])
        assert rc != 0 and err.strip(), (cut, rc)
