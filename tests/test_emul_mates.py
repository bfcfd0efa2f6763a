"""CPU-side check of the `-m` (fix-mate-overlaps) device logic: sambamba_b200/csrc/mates.cuh compiled for the host
(tests/emul/emul_mates.cpp) and run thread by thread over the same record SoA k2_decode produces, against the
oracle's faithful column sweep (oracle/depth_oracle.c, depth.d:319-399,495-556,760-845) and its closed form.
No GPU needed."""
import ctypes as C
import os
import random
import struct

import numpy as np
import pytest

import helpers
from helpers import GOLDEN, ROOT

U64P, U32P, I64P, I32P, U8P = (C.POINTER(t) for t in (C.c_uint64, C.c_uint32, C.c_int64, C.c_int32, C.c_uint8))


@pytest.fixture(scope="module")
def em():
    L = C.CDLL(os.path.join(ROOT, "tests", "emul", "libemul_mates.so"))
    L.emul_mates.restype = C.c_int
    return L


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None and a.size else C.cast(None, t)


class Soa:
    """The record table of a BAM file, field for field what k2_decode writes (kernels.cuh RecordSoA)."""

    def __init__(self, path, mapq_gt=0, flag_reject=0x600):
        self.u = helpers.oracle_inflate(path)
        first, refs = helpers.header_first_record_offset(self.u)
        self.refs = refs
        b = self.u.tobytes()
        l_text = struct.unpack_from("<i", b, 4)[0]
        text = b[8:8 + l_text].decode(errors="replace")
        self.samples, rg2s = [], {}
        for line in text.split("\n"):
            if line.startswith("@RG"):
                f = dict(x.split(":", 1) for x in line.split("\t")[1:] if ":" in x)
                sm = f.get("SM", "")
                if sm not in self.samples:
                    self.samples.append(sm)
                rg2s[f.get("ID", "")] = self.samples.index(sm)
        if not self.samples:
            self.samples = ["*"]
        lin0, t = [], 0
        for _, ln in refs:
            lin0.append(t)
            t += ln
        self.lin0, self.total = lin0, t
        rows = []
        self.recs = []          # python-side view for the plain scatter: (start, span, sample, cigar, seq nibbles, qual, pass)
        o = first
        while o + 4 <= len(b):
            bs = struct.unpack_from("<i", b, o)[0]
            if o + 4 + bs > len(b):
                break
            ref, pos, bmn, fnc, lseq = struct.unpack_from("<iiIIi", b, o + 4)
            l_name, mapq, flag, ncig = bmn & 0xFF, (bmn >> 8) & 0xFF, fnc >> 16, fnc & 0xFFFF
            cg0 = o + 36 + l_name
            cig = [struct.unpack_from("<I", b, cg0 + 4 * k)[0] for k in range(ncig)]
            span = sum(c >> 4 for c in cig if (c & 15) in (0, 2, 3, 7, 8))
            placed = 0 <= ref < len(refs) and pos >= 0
            ok = placed and mapq > mapq_gt and not (flag & flag_reject) and not (flag & 4) and span > 0
            start = lin0[ref] + pos if placed else 0xFFFFFFFFFFFFFFFE
            span_eff = 0
            if ok:
                span_eff = min(span, max(0, refs[ref][1] - pos))
                ok = span_eff > 0
            sample = 0
            if rg2s and ok:
                a = cg0 + 4 * ncig + (lseq + 1) // 2 + lseq
                end = o + 4 + bs
                while a + 3 <= end:
                    tag, ty = b[a:a + 2], chr(b[a + 2])
                    v = a + 3
                    if tag == b"RG" and ty == "Z":
                        z = b.index(b"\0", v)
                        sample = rg2s[b[v:z].decode()]
                        break
                    if ty in "AcC":
                        n = 1
                    elif ty in "sS":
                        n = 2
                    elif ty in "iIf":
                        n = 4
                    elif ty in "ZH":
                        n = b.index(b"\0", v) - v + 1
                    elif ty == "B":
                        st, cnt = chr(b[v]), struct.unpack_from("<I", b, v + 1)[0]
                        n = 5 + cnt * (1 if st in "cC" else 2 if st in "sS" else 4)
                    else:
                        break
                    a = v + n
            rows.append((start, span_eff, (flag << 16) | (mapq << 8) | (sample << 2) | (1 if ok else 0) | (2 if ok and span_eff > 1024 else 0), o + 4, (ncig << 8) | l_name, lseq))
            sq0 = cg0 + 4 * ncig
            self.recs.append((start, span_eff, sample, cig, sq0, sq0 + (lseq + 1) // 2, max(lseq, 0), ok))
            o += 4 + bs
        R = len(rows)
        self.R = R
        self.start = np.array([r[0] for r in rows], np.uint64)
        self.span = np.array([r[1] for r in rows], np.uint32)
        self.meta = np.array([r[2] for r in rows], np.uint32)
        self.off = np.array([r[3] for r in rows], np.int64)
        self.ncl = np.array([r[4] for r in rows], np.uint32)
        self.lseq = np.array([r[5] for r in rows], np.int32)

    def plain_counts(self, S, minq=0):
        """Per-read scatter (what K3 leaves in the counter planes): [S, 7, total]."""
        out = np.zeros((S, 7, self.total), np.uint32)
        b = self.u
        nt5 = {1: 0, 2: 1, 4: 2, 8: 3}
        for start, span, sample, cig, sq, ql, lseq, ok in self.recs:
            if not ok:
                continue
            s = sample if S > 1 else 0
            rp = qp = 0
            for c in cig:
                ln, op = c >> 4, c & 15
                if op in (0, 7, 8):
                    for k in range(ln):
                        if rp + k < span and qp + k < lseq and b[ql + qp + k] >= minq:
                            q = qp + k
                            nib = (b[sq + (q >> 1)] & 15) if q & 1 else (b[sq + (q >> 1)] >> 4)
                            out[s, nt5.get(int(nib), 4), start + rp + k] += 1
                    rp += ln
                    qp += ln
                elif op in (2, 3):
                    lo, hi = rp, min(rp + ln, span)
                    if hi > lo:
                        out[s, 5 if op == 2 else 6, start + lo:start + hi] += 1
                    rp += ln
                elif op in (1, 4):
                    qp += ln
        return out

    def read_hits(self, a, b, minq=0):
        """countRead (depth.d:661-669) of every passing read against [a, b): per-sample number of reads with >= 1 counted base."""
        n = {}
        buf = self.u
        for start, span, sample, cig, sq, ql, lseq, ok in self.recs:
            if not ok or start >= b or start + span <= a:
                continue
            rp = qp = 0
            hit = False
            for c in cig:
                ln, op = c >> 4, c & 15
                if op in (0, 7, 8):
                    for k in range(ln):
                        g = start + rp + k
                        if a <= g < b and rp + k < span and qp + k < lseq and buf[ql + qp + k] >= minq:
                            hit = True
                            break
                    rp += ln
                    qp += ln
                elif op in (2, 3):
                    rp += ln
                elif op in (1, 4):
                    qp += ln
                if hit:
                    break
            if hit:
                n[sample] = n.get(sample, 0) + 1
        return n


def run_emul(em, soa, counts, S, minq=0, flt=None, segs=None, n_samples_out=1, force_general=0, order=0, cnt_base=0, seg_u=None, seg_qmin=None, ext_max=0):
    """counts: [S, 7, total] (modified in place).  flt: merged linear (start, end) list.  segs: linear (start, end) list in
    output order.  Returns (err, stats, seg_reads_delta [n_samples_out, n_seg], seg_mbases)."""
    fs = np.array([x[0] for x in (flt or [])], np.uint64)
    fe = np.array([x[1] for x in (flt or [])], np.uint64)
    n_seg = len(segs or [])
    order_idx = sorted(range(n_seg), key=lambda i: (segs[i][0], i))
    ss = np.array([segs[i][0] for i in order_idx], np.uint64)
    se = np.array([segs[i][1] for i in order_idx], np.uint64)
    pm = np.maximum.accumulate(se) if n_seg else np.zeros(0, np.uint64)
    sid = np.array(order_idx, np.uint32)
    su = np.array([seg_u[i] for i in order_idx], np.uint64) if seg_u is not None else None
    sq = np.array([seg_qmin[i] for i in order_idx], np.uint64) if seg_qmin is not None else None
    sreads = np.zeros((n_samples_out, max(n_seg, 1)), np.uint32)
    smb = np.zeros((n_samples_out, max(n_seg, 1)), np.uint32)
    err = (C.c_int * 2)()
    stat = (C.c_ulonglong * 3)()
    u = soa.u
    rc = em.emul_mates(_ptr(u, U8P), C.c_uint32(soa.R), _ptr(soa.start, U64P), _ptr(soa.span, U32P), _ptr(soa.meta, U32P), _ptr(soa.off, I64P),
                       _ptr(soa.ncl, U32P), _ptr(soa.lseq, I32P), _ptr(fs, U64P), _ptr(fe, U64P), C.c_uint32(len(fs)),
                       _ptr(counts, U32P), C.c_uint64(cnt_base), C.c_uint64(counts.shape[2]), C.c_uint32(S), C.c_uint32(minq),
                       _ptr(ss, U64P), _ptr(se, U64P), _ptr(pm, U64P), _ptr(sid, U32P), C.c_uint32(n_seg), _ptr(sreads, U32P), _ptr(smb, U32P),
                       C.c_uint32(n_samples_out), C.c_int(force_general), C.c_int(order), err, stat,
                       _ptr(su, U64P) if seg_u is not None else C.cast(None, U64P), _ptr(sq, U64P) if seg_qmin is not None else C.cast(None, U64P), C.c_uint64(ext_max))
    return rc, list(stat), sreads.view(np.int32)[:, :n_seg], smb.view(np.int32)[:, :n_seg], (err[0], err[1])      # corrections are signed (the library adds them modulo 2^32)


# ------------------------------------------------------------------------------------------------ inputs
def random_cigar(rnd, qlen, allow_n=True):
    """A CIGAR consuming qlen query bases: optional soft clips, M runs separated by I / D / N (never leading or trailing)."""
    ops = []
    left = qlen
    if rnd.random() < 0.15:
        s = rnd.randint(1, 5)
        ops.append((s, 4))
        left -= s
    tail = rnd.randint(1, 5) if rnd.random() < 0.15 else 0
    left -= tail
    n_m = rnd.choice([1, 1, 1, 2, 2, 3])
    cuts = sorted(rnd.sample(range(1, left), n_m - 1)) if n_m > 1 else []
    ms = [b - a for a, b in zip([0] + cuts, cuts + [left])]
    for i, m in enumerate(ms):
        if i:
            k = rnd.random()
            if k < 0.35:
                ins = min(rnd.randint(1, 4), m - 1) if m > 1 else 0
                if ins:
                    ops.append((ins, 1))
                    m -= ins
            elif k < 0.75 or not allow_n:
                ops.append((rnd.randint(1, 6), 2))
            else:
                ops.append((rnd.randint(5, 40), 3))
        ops.append((m, 0))
    if tail:
        ops.append((tail, 4))
    return ops


def make_pairs_bam(path, seed, n_frag=400, refs=(("c1", 4000), ("c2", 2500)), rg=None, triples=0.0, same_start=0.05, chains=0.0, pile=0):
    rnd = random.Random(seed)
    reads, quals, tags = [], [], []
    rgs = [r[0] for r in rg] if rg else None

    def one(ref, pos, name, flag, rgid):
        qlen = rnd.randint(30, 90)
        cig = random_cigar(rnd, qlen)
        seq = "".join(rnd.choice("ACGTACGTACGTN") for _ in range(qlen))
        mapq = rnd.choice([0, 3, 20, 20, 60, 60, 60])
        span = sum(l for l, op in cig if op in (0, 2, 3))
        pos = max(0, min(pos, refs[ref][1] - span - 1))
        reads.append((ref, pos, mapq, flag, cig, seq, name))
        quals.append([rnd.choice([2, 10, 20, 20, 30, 30, 37, 41]) for _ in range(qlen)])
        tags.append((b"RGZ" + rgid.encode() + b"\0") if rgid else b"")

    for i in range(n_frag):
        ref = rnd.randrange(len(refs))
        pos = rnd.randint(0, refs[ref][1] - 200)
        name = "frag%05d" % i
        rgid = rnd.choice(rgs) if rgs else None
        k = rnd.random()
        flag_extra = rnd.choice([0, 0, 0, 0, 0, 0x400, 0x200])
        if k < 0.15:
            one(ref, pos, name, 0, rgid)                                   # single read
            continue
        one(ref, pos, name, 0x41 | flag_extra, rgid)
        delta = 0 if rnd.random() < same_start else rnd.randint(0, 110)
        one(ref, pos + delta, name, 0x81 | (0 if rnd.random() < 0.9 else rnd.choice([0x400, 0x200])), rgid if rnd.random() < 0.97 or not rgs else rnd.choice(rgs))
        if rnd.random() < triples:
            for _ in range(rnd.choice([1, 1, 2])):
                one(ref, pos + rnd.randint(0, 120), name, 0x800 | 0x41, rgid)
        if rnd.random() < chains:                  # a long chain of supplementary alignments of one name, a few of them present in any column
            x = pos
            for _ in range(rnd.randint(8, 30)):
                x += rnd.randint(25, 60)
                if x > refs[ref][1] - 250:
                    break                          # (one() clamps positions at the reference end: the chain would pile up there)
                one(ref, x, name, 0x800 | 0x41, rgid)
    for j in range(pile):                          # `pile` reads of one name over one position
        one(0, 1000 + j, "pile", 0x800 | 0x41, rgs[0] if rgs else None)
    order = sorted(range(len(reads)), key=lambda j: (reads[j][0], reads[j][1]))
    helpers.write_bam(path, list(refs), [reads[j] for j in order], rg=rg, quals=[quals[j] for j in order], tags=[tags[j] for j in order], block=rnd.choice([0xFF00, 3000]))
    return path


def sweep_base_counts(path, total_by_ref, extra=()):
    """Counters of the oracle's faithful sweep (`depth base -m -c 0 --combined` text) -> [6, total] (A,C,G,T,DEL,REFSKIP) + COV."""
    rc, out, err = helpers.oracle_cli(["base", "-m", "-c", "0", "--combined", *extra, path])
    assert rc == 0, err
    names = [n for n, _ in total_by_ref]
    lin0, t = {}, 0
    for n, ln in total_by_ref:
        lin0[n] = t
        t += ln
    cov = np.zeros(t, np.uint32)
    pl = np.zeros((6, t), np.uint32)
    for line in out.decode().split("\n")[1:]:
        if not line:
            continue
        f = line.split("\t")
        g = lin0[f[0]] + int(f[1])
        cov[g] = int(f[2])
        pl[:, g] = [int(x) for x in f[3:9]]
    assert names
    return cov, pl


def planes6(c7):
    return np.stack([c7[0], c7[1], c7[2], c7[3], c7[5], c7[6]])


# ------------------------------------------------------------------------------------------------ tests
@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("minq", [0, 20])
def test_pairs_base_mode_matches_sweep(em, tmp_path, seed, minq):
    p = make_pairs_bam(str(tmp_path / f"pairs{seed}.bam"), seed)
    soa = Soa(p)
    want_cov, want6 = sweep_base_counts(p, soa.refs, extra=("-q", str(minq)) if minq else ())
    plain = soa.plain_counts(1, minq)
    ref_plain, _ = helpers.oracle_counts(p, min_bq=minq)
    assert np.array_equal(plain[0], ref_plain), "the test's own scatter disagrees with the oracle's closed form"
    for order in (0, 1, 2):
        for general in (0, 1):
            c = plain.copy()
            rc, stat, _, _, err = run_emul(em, soa, c, 1, minq=minq, force_general=general, order=order)
            assert rc == 0, err
            assert np.array_equal(c[0].sum(axis=0), want_cov), (order, general)
            assert np.array_equal(planes6(c[0]), want6), (order, general)
            if not general:
                assert stat[0] > 50 and stat[1] > 1000      # the generator does make overlapping pairs
    closed, _ = helpers.oracle_counts_fix_mates(p, min_bq=minq)
    c = plain.copy()
    run_emul(em, soa, c, 1, minq=minq)
    assert np.array_equal(c[0], closed)


def test_fixtures_base_mode(em):
    for name in ("issue_204.bam", "mate_overlaps_1_3M_4M.bam", "issue_193.bam"):
        p = os.path.join(GOLDEN, name)
        soa = Soa(p)
        win = helpers.interesting_window(p)
        closed, npc = helpers.oracle_counts_fix_mates(p, window=win)
        plain, _ = helpers.oracle_counts(p, window=win)
        c = np.ascontiguousarray(plain[None])
        rc, stat, _, _, err = run_emul(em, soa, c, 1, cnt_base=win[0])
        assert rc == 0, (name, err)
        assert np.array_equal(c[0], closed), name
        assert stat[1] == npc, name


def test_groups_of_three_and_more(em, tmp_path):
    """Supplementary alignments sharing a name: the state machine path (past reads pairing again count twice; a flagged
    read left over as the last entry of the column's sorted array keeps its state, depth.d:380-384)."""
    done = 0
    for seed in range(10, 30):
        p = make_pairs_bam(str(tmp_path / f"tri{seed}.bam"), seed, n_frag=120, triples=0.5)
        soa = Soa(p)
        plain = soa.plain_counts(1)
        c = plain.copy()
        rc, stat, _, _, err = run_emul(em, soa, c, 1, order=seed % 3)
        assert rc == 0, err
        want_cov, want6 = sweep_base_counts(p, soa.refs)
        assert stat[2] > 0
        assert np.array_equal(planes6(c[0]), want6) and np.array_equal(c[0].sum(axis=0), want_cov), seed
        done += 1
    assert done == 20


def test_long_chains_of_one_name(em, tmp_path):
    """A component (chain of overlapping same-name reads) of any length: only the number of reads of one name in ONE column is bounded."""
    for seed in range(300, 310):
        p = make_pairs_bam(str(tmp_path / f"ch{seed}.bam"), seed, n_frag=60, triples=0.3, chains=0.3)
        soa = Soa(p)
        c = soa.plain_counts(1)
        rc, stat, _, _, err = run_emul(em, soa, c, 1, order=seed % 3)
        assert rc == 0, err
        want_cov, want6 = sweep_base_counts(p, soa.refs)
        assert stat[2] > 0
        assert np.array_equal(planes6(c[0]), want6) and np.array_equal(c[0].sum(axis=0), want_cov), seed
        closed, npc = helpers.oracle_counts_fix_mates(p)
        assert np.array_equal(c[0], closed) and stat[1] == npc, seed
    # region mode over the same kind of file
    for seed in (311, 312, 313):
        p = make_pairs_bam(str(tmp_path / f"chr{seed}.bam"), seed, n_frag=60, triples=0.3, chains=0.3)
        rnd = random.Random(seed)
        regs = []
        for ref, ln in ((0, 4000), (1, 2500)):
            x = rnd.randint(0, 60)
            while x < ln - 50:
                w = rnd.choice([1, 7, 30, 90, 200, 600])
                regs.append((ref, x, min(x + w, ln)))
                x += w + rnd.choice([0, 0, 1, 5, 40, 300])
        st = check_regions(em, p, regs, [1, 4], rnd.choice([0, 20]), tmp_path)
        assert st[2] > 0


def test_more_reads_of_one_name_in_a_column_than_the_walk_holds(em, tmp_path):
    """Eight reads of one name over one position are replayed; a ninth is refused (never silently different)."""
    p = make_pairs_bam(str(tmp_path / "p8.bam"), 5, n_frag=20, pile=8)
    soa = Soa(p)
    c = soa.plain_counts(1)
    rc, stat, _, _, err = run_emul(em, soa, c, 1)
    assert rc == 0, err
    want_cov, want6 = sweep_base_counts(p, soa.refs)
    assert np.array_equal(planes6(c[0]), want6) and np.array_equal(c[0].sum(axis=0), want_cov)
    p = make_pairs_bam(str(tmp_path / "p9.bam"), 5, n_frag=20, pile=9)
    soa = Soa(p)
    rc, stat, _, _, err = run_emul(em, soa, soa.plain_counts(1), 1)
    assert err[0] == 1, (rc, err)


def test_multi_sample_pairs(em, tmp_path):
    rg = [("g1", "S1"), ("g2", "S2"), ("g3", "S1")]
    p = make_pairs_bam(str(tmp_path / "ms.bam"), 7, rg=rg)
    soa = Soa(p)
    assert soa.samples == ["S1", "S2"]
    plain = soa.plain_counts(2)
    c = plain.copy()
    rc, stat, _, _, err = run_emul(em, soa, c, 2, n_samples_out=2)
    assert rc == 0, err
    rc2, out, e2 = helpers.oracle_cli(["base", "-m", "-c", "0", p])
    assert rc2 == 0, e2
    lin0 = dict(zip([n for n, _ in soa.refs], soa.lin0))
    seen = 0
    for line in out.decode().split("\n")[1:]:
        if not line:
            continue
        f = line.split("\t")
        g, s = lin0[f[0]] + int(f[1]), soa.samples.index(f[9])
        got = c[s, :, g]
        assert [int(x) for x in f[3:9]] == [int(got[0]), int(got[1]), int(got[2]), int(got[3]), int(got[5]), int(got[6])] and int(f[2]) == int(got.sum()), line
        seen += 1
    assert seen > 1000
    # --combined: one counter set, mates of different samples still do not pair
    c1 = soa.plain_counts(1)
    rc, _, _, _, err = run_emul(em, soa, c1, 1)
    assert rc == 0
    want_cov, want6 = sweep_base_counts(p, soa.refs)
    assert np.array_equal(planes6(c1[0]), want6)


def _fmt_g(v):
    return "%g" % float(np.float32(v))


def region_rows(soa, counts, regs, thr, sreads, smb, minq, samples_out):
    """Text rows `depth region` prints for regs [(ref, start, end)], from the fixed counters and the pair corrections."""
    rows = []
    for i, (ref, a, b) in enumerate(regs):
        la, lb = soa.lin0[ref] + a, soa.lin0[ref] + min(b, soa.refs[ref][1])
        hits = soa.read_hits(la, lb, minq)
        for s in range(samples_out):
            pl = counts[s]
            n_bases = int(pl[:5, la:lb].sum()) + int(smb[s, i])
            if samples_out > 1:
                n_reads = hits.get(s, 0) + int(sreads[s, i])
            else:
                n_reads = sum(hits.values()) + int(sreads[0, i])
            cov = pl[:, la:lb].sum(axis=0)
            ln = np.float32(b - a)
            f = [soa.refs[ref][0], str(a), str(b), str(n_reads), _fmt_g(np.float32(n_bases) / ln)]
            for t in thr:
                f.append("100" if t == 0 else _fmt_g(np.float32(int((cov >= t).sum())) * np.float32(100) / ln))
            if samples_out > 1 or True:
                f.append(soa.samples[s] if samples_out > 1 else None)
            rows.append(f)
    return rows


def merged(regs, soa):
    lin = sorted((soa.lin0[r] + a, soa.lin0[r] + min(b, soa.refs[r][1])) for r, a, b in regs)
    out = []
    for a, b in lin:
        if out and out[-1][1] >= a:
            out[-1][1] = max(out[-1][1], b)
        else:
            out.append([a, b])
    return [tuple(x) for x in out]


def check_regions(em, p, regs, thr, minq, tmp_path, combined=True, force_general=0):
    soa = Soa(p)
    S = 1 if combined else len(soa.samples)
    bed = str(tmp_path / "r.bed")
    with open(bed, "w") as f:
        for r, a, b in regs:
            f.write(f"{soa.refs[r][0]}\t{a}\t{b}\n")
    args = ["region", "-L", bed, "-m"] + [x for t in thr for x in ("-T", str(t))] + (["-q", str(minq)] if minq else []) + (["--combined"] if combined else []) + [p]
    rc, out, err = helpers.oracle_cli(args)
    assert rc == 0, err
    want = [l.split("\t") for l in out.decode().split("\n")[1:] if l]
    flt = merged(regs, soa)
    segs = [(soa.lin0[r] + a, soa.lin0[r] + min(b, soa.refs[r][1])) for r, a, b in regs]
    counts = soa.plain_counts(S, minq)
    rc, stat, sreads, smb, e = run_emul(em, soa, counts, S, minq=minq, flt=flt, segs=segs, n_samples_out=S, force_general=force_general, order=2 if force_general else 0)
    assert rc == 0, e
    got = region_rows(soa, counts, regs, thr, sreads, smb, minq, S)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        g = [x for x in g if x is not None]
        assert g == w, (g, w)
    return stat


@pytest.mark.parametrize("seed", [5, 6, 7, 8])
def test_pairs_region_mode_matches_sweep(em, tmp_path, seed):
    p = make_pairs_bam(str(tmp_path / f"reg{seed}.bam"), seed, n_frag=500)
    rnd = random.Random(seed)
    # sorted, disjoint regions (NonOverlappingRegionStatsCollector), some adjacent, some tiny, some starting inside pairs
    regs = []
    for ref, ln in ((0, 4000), (1, 2500)):
        x = rnd.randint(0, 100)
        while x < ln - 50:
            w = rnd.choice([1, 7, 30, 90, 200, 400])
            regs.append((ref, x, min(x + w, ln)))
            x += w + rnd.choice([0, 0, 1, 5, 40, 150])
    for minq in (0, 20):
        stat = check_regions(em, p, regs, [0, 1, 3, 8], minq, tmp_path)
        assert stat is not None and stat[0] > 30


def test_overlapping_regions_general_collector(em, tmp_path):
    p = make_pairs_bam(str(tmp_path / "gen.bam"), 11, n_frag=500)
    rnd = random.Random(11)
    regs = [(rnd.randrange(2), a, a + rnd.choice([5, 60, 300])) for a in (rnd.randint(0, 2000) for _ in range(40))]
    regs = [(r, a, min(b, 4000 if r == 0 else 2500)) for r, a, b in regs]
    check_regions(em, p, regs, [1, 5], 0, tmp_path)


def test_golden_issue_204_region(em, tmp_path):
    """The reference's only region-mode golden vector (test_suite.sh:156-162) through the device logic."""
    p = os.path.join(GOLDEN, "issue_204.bam")
    soa = Soa(p)
    ref = [n for n, _ in soa.refs].index("2")
    a, b = 166868600 - 1, 166868813
    flt = [(soa.lin0[ref] + a, soa.lin0[ref] + b)]
    win = (flt[0][0] - 2000, flt[0][1] + 2000)
    # counters over a window (the genome is human sized): scatter only what lies in it
    plain, _ = helpers.oracle_counts(p, window=win)
    c = np.ascontiguousarray(plain[None])
    rc, stat, sreads, smb, err = run_emul(em, soa, c, 1, flt=flt, segs=flt, cnt_base=win[0])
    assert rc == 0, err
    la, lb = flt[0][0] - win[0], flt[0][1] - win[0]
    n_bases = int(c[0, :5, la:lb].sum()) + int(smb[0, 0])
    n_reads = sum(soa.read_hits(flt[0][0], flt[0][1]).values()) + int(sreads[0, 0])
    cov = c[0, :, la:lb].sum(axis=0)
    ln = np.float32(b - a)
    row = ["2", str(a), str(b), str(n_reads), _fmt_g(np.float32(n_bases) / ln)] + [_fmt_g(np.float32(int((cov >= t).sum())) * np.float32(100) / ln) for t in (15, 20, 25)]
    want = open(os.path.join(GOLDEN, "issue_204_expected_output.txt")).read().split("\n")[1].split("\t")
    assert row == want[:len(row)], (row, want)


def test_generated_pairs_with_real_cigar_mix(em, tmp_path):
    """tools/bamgen --pairs: the synthetic mix of SURVEY 8d (150 bp reads, I/D/S/N CIGARs) with mates sharing names."""
    p = helpers.gen_bam(str(tmp_path / "gp.bam"), "--preset", "tiny", "-n", 6000, "--pairs", 6, "-t", 2)
    soa = Soa(p)
    for minq in (0, 25):
        want, npc = helpers.oracle_counts_fix_mates(p, min_bq=minq)
        plain, _ = helpers.oracle_counts(p, min_bq=minq)
        c = np.ascontiguousarray(plain[None])
        rc, stat, _, _, err = run_emul(em, soa, c, 1, minq=minq, order=2)
        assert rc == 0, err
        assert stat[1] == npc and npc > 10000
        assert np.array_equal(c[0], want)
    regs = [(0, 100, 900), (0, 1000, 1200), (0, 1200, 1207), (1, 10, 20), (2, 5, 40000)]
    check_regions(em, p, regs, [3, 10], 0, tmp_path)


@pytest.mark.parametrize("seed,w", [(31, 100), (32, 333), (33, 64), (34, 1000)])
def test_pairs_window_mode_without_overlap(em, tmp_path, seed, w):
    """`depth window -w W -m` with --overlap 0: one ring slot, every column lies in the window being filled, windows
    follow each other like sorted adjacent regions (first_occ is set again when a window is finished, depth.d:962-972)."""
    p = make_pairs_bam(str(tmp_path / f"win{seed}.bam"), seed, n_frag=500)
    soa = Soa(p)
    for minq in (0, 20):
        args = ["window", "-w", str(w), "-m", "--combined", "-T", "2", "-T", "6"] + (["-q", str(minq)] if minq else []) + [p]
        rc, out, err = helpers.oracle_cli(args)
        assert rc == 0, err
        want = [l.split("\t") for l in out.decode().split("\n")[1:] if l]
        regs = []
        for ref, (_, ln) in enumerate(soa.refs):
            regs += [(ref, k * w, (k + 1) * w) for k in range(ln // w)]
        extra = [(ref, (ln // w) * w, (ln // w + 1) * w) for ref, (_, ln) in enumerate(soa.refs)]      # the partial window at the end of a reference is a segment too (never printed)
        allr = regs + extra
        segs = [(soa.lin0[r] + a, soa.lin0[r] + min(b, soa.refs[r][1])) for r, a, b in allr]
        counts = soa.plain_counts(1, minq)
        rc, stat, sreads, smb, e = run_emul(em, soa, counts, 1, minq=minq, segs=segs)
        assert rc == 0, e
        got = region_rows(soa, counts, regs, [2, 6], sreads, smb, minq, 1)
        got = [[x for x in g if x is not None] for g in got]
        assert got == want


@pytest.mark.parametrize("seed", [5, 8])
def test_region_mode_state_machine_path_on_pairs(em, tmp_path, seed):
    """The per-name replay of PerRegionPrinter.push (used for 3+ reads of a name) must agree with the pair closed form."""
    p = make_pairs_bam(str(tmp_path / f"rg{seed}.bam"), seed, n_frag=400)
    rnd = random.Random(seed)
    regs = []
    for ref, ln in ((0, 4000), (1, 2500)):
        x = rnd.randint(0, 60)
        while x < ln - 50:
            w = rnd.choice([1, 7, 30, 90, 200])
            regs.append((ref, x, min(x + w, ln)))
            x += w + rnd.choice([0, 0, 1, 5, 40])
    for minq in (0, 20):
        assert check_regions(em, p, regs, [1, 4], minq, tmp_path, force_general=1) is not None


def test_region_mode_groups_of_three_and_more(em, tmp_path):
    done = 0
    for seed in range(40, 64):
        p = make_pairs_bam(str(tmp_path / f"trg{seed}.bam"), seed, n_frag=150, triples=0.5)
        rnd = random.Random(seed)
        regs = []
        for ref, ln in ((0, 4000), (1, 2500)):
            x = rnd.randint(0, 60)
            while x < ln - 50:
                w = rnd.choice([1, 7, 30, 90, 200, 600])
                regs.append((ref, x, min(x + w, ln)))
                x += w + rnd.choice([0, 0, 1, 5, 40, 300])
        st = check_regions(em, p, regs, [1, 4], rnd.choice([0, 20]), tmp_path)
        assert st[2] > 0
        done += 1
    assert done == 24
    # overlapping regions too (GeneralRegionStatsCollector)
    p = make_pairs_bam(str(tmp_path / "trgo.bam"), 77, n_frag=150, triples=0.5)
    rnd = random.Random(77)
    regs = [(rnd.randrange(2), a, a + rnd.choice([5, 60, 300])) for a in (rnd.randint(0, 2000) for _ in range(30))]
    check_regions(em, p, regs, [2], 0, tmp_path)


def window_segments(soa, w, overlap):
    """The slots bdepth_run_windows defines (bdepth.cu): per reference max(n_full + nslot, n_empty) windows k*step, with the
    early-update extent of reused ring slots and the first-occurrence quirk of reference 0's first slots."""
    step = w - overlap
    nslot = (w + step - 1) // step
    ext = nslot * step - w
    segs, n_full = [], []
    for r, (_, L) in enumerate(soa.refs):
        nf = (L - w) // step + 1 if L >= w else 0
        n_full.append(nf)
        for k in range(max(nf + nslot, L // step)):
            segs.append(dict(ref=r, start=k * step, end=k * step + w, cov_ext=ext if k >= nslot else 0, qmin=k * step if (r == 0 and 1 <= k < nslot) else 0))
    return segs, n_full, step


def window_rows(soa, counts, segs, n_full, thr, sreads, smb, minq):
    """Rows of `depth window` for references that all have reads, as run_segments + the reducers compute them: bases over
    the window, thresholds from the slot's first updated column, reads starting at/after qmin for the quirk slots."""
    rows, i0 = [], 0
    per_ref = {}
    for i, sd in enumerate(segs):
        per_ref.setdefault(sd["ref"], []).append(i)
    for r, idxs in per_ref.items():
        L, l0 = soa.refs[r][1], soa.lin0[r]
        for i in idxs[:n_full[r]]:
            sd = segs[i]
            a, b = l0 + min(sd["start"], L), l0 + min(sd["end"], L)
            ac = l0 + min(sd["start"] - min(sd["cov_ext"], sd["start"]), L)
            if sd["qmin"]:
                nb, nr = 0, 0
                for start, span, sample, cig, sq, ql, lseq, ok in soa.recs:
                    if not ok or start < l0 + sd["qmin"] or start >= b or start + span <= a:
                        continue
                    rp = qp = n = 0
                    for c in cig:
                        ln, op = c >> 4, c & 15
                        if op in (0, 7, 8):
                            for k in range(ln):
                                g = start + rp + k
                                if a <= g < b and rp + k < span and qp + k < lseq and soa.u[ql + qp + k] >= minq:
                                    n += 1
                            rp += ln
                            qp += ln
                        elif op in (2, 3):
                            rp += ln
                        elif op in (1, 4):
                            qp += ln
                    nb += n
                    nr += n > 0
            else:
                nb = int(counts[0, :5, a:b].sum())
                nr = sum(soa.read_hits(a, b, minq).values())
            nb = (nb + int(smb[0, i])) & 0xFFFFFFFF
            nr = (nr + int(sreads[0, i])) & 0xFFFFFFFF
            cov = counts[0, :, ac:b].sum(axis=0)
            ln = np.float32(sd["end"] - sd["start"])
            f = [soa.refs[r][0], str(sd["start"]), str(sd["end"]), str(nr), _fmt_g(np.float32(nb) / ln)]
            for t in thr:
                f.append("100" if t == 0 else _fmt_g(np.float32(int((cov >= t).sum())) * np.float32(100) / ln))
            rows.append(f)
    return rows


@pytest.mark.parametrize("seed,w,overlap", [(51, 100, 50), (52, 100, 30), (53, 90, 80), (54, 333, 100), (55, 64, 63)])
def test_overlapping_windows(em, tmp_path, seed, w, overlap):
    """`depth window --overlap` with and without -m.  Without -m this checks the test's own statement of what the library
    computes for ring slots (early threshold collection, first-occurrence quirk); with -m the per-name replay adds its terms."""
    p = make_pairs_bam(str(tmp_path / f"ow{seed}.bam"), seed, n_frag=300, refs=(("c1", 2200), ("c2", 1300)), triples=0.2 if seed % 2 else 0.0)
    soa = Soa(p)
    segs, n_full, step = window_segments(soa, w, overlap)
    lin = [(soa.lin0[s["ref"]] + min(s["start"], soa.refs[s["ref"]][1]), soa.lin0[s["ref"]] + min(s["end"], soa.refs[s["ref"]][1])) for s in segs]
    us = [min(a, soa.lin0[s["ref"]] + min(s["start"] - min(s["cov_ext"], s["start"]), soa.refs[s["ref"]][1])) for s, (a, _) in zip(segs, lin)]
    qm = [soa.lin0[s["ref"]] + s["qmin"] if s["qmin"] else 0 for s in segs]
    for minq in (0, 20):
        base_args = ["window", "-w", str(w), "--overlap", str(overlap), "--combined", "-T", "2", "-T", "5"] + (["-q", str(minq)] if minq else [])
        # without -m: zero corrections
        rc, out, err = helpers.oracle_cli(base_args + [p])
        assert rc == 0, err
        want = [l.split("\t") for l in out.decode().split("\n")[1:] if l]
        counts = soa.plain_counts(1, minq)
        z = np.zeros((1, len(segs)), np.int64)
        assert window_rows(soa, counts, segs, n_full, [2, 5], z, z, minq) == want
        # with -m
        rc, out, err = helpers.oracle_cli(base_args + ["-m", p])
        assert rc == 0, err
        want = [l.split("\t") for l in out.decode().split("\n")[1:] if l]
        rc, stat, sreads, smb, e = run_emul(em, soa, counts, 1, minq=minq, segs=lin, seg_u=us, seg_qmin=qm, ext_max=max(a - u for (a, _), u in zip(lin, us)), order=seed % 3)
        assert rc == 0, e
        got = window_rows(soa, counts, segs, n_full, [2, 5], sreads, smb, minq)
        assert got == want
