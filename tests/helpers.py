"""Shared test helpers: oracle bindings (TEST INFRASTRUCTURE), synthetic inputs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_EXE = os.path.join(ROOT, "oracle", "_build", "depth_oracle")
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
BAMGEN = os.path.join(ROOT, "tools", "_build", "bamgen")
CLI = os.path.join(ROOT, "sambamba_b200", "_build", "sambamba-depth-b200")

_orc = None


class ScatterStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_records", "n_pass", "n_blocks", "ulen", "clen", "covered", "min_lin", "max_lin")] + \
               [("t_inflate", C.c_double), ("t_scan", C.c_double)]


def oracle():
    global _orc
    if _orc is None:
        L = C.CDLL(ORACLE_LIB)
        L.oracle_inflate_file.restype = C.c_int64
        L.oracle_inflate_file.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64]
        L.oracle_base_counts.argtypes = [C.c_char_p, C.c_int, C.c_uint, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(ScatterStats)]
        L.oracle_base_counts_fix_mates.argtypes = [C.c_char_p, C.c_int, C.c_uint, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
        L.oracle_bam_info.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_segment_stats.argtypes = [C.c_char_p, C.c_int, C.c_uint, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _orc = L
    return _orc


def oracle_inflate(path):
    L = oracle()
    n = L.oracle_inflate_file(path.encode(), None, 0)
    assert n >= 0, L.oracle_last_error()
    buf = np.zeros(max(1, n), np.uint8)
    assert L.oracle_inflate_file(path.encode(), buf.ctypes.data_as(C.c_void_p), n) == n
    return buf[:n]


def oracle_info(path):
    L = oracle()
    nref, tot, ulen, nblk = C.c_int(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    assert L.oracle_bam_info(path.encode(), C.byref(nref), C.byref(tot), C.byref(ulen), C.byref(nblk)) == 0
    return nref.value, tot.value, ulen.value, nblk.value


def oracle_scan(path, mapq_gt=0, flag_reject=0x600):
    """Record counts and the linear-coordinate extent [min_lin, max_lin) of the passing reads."""
    L = oracle()
    st = ScatterStats()
    rc = L.oracle_base_counts(path.encode(), mapq_gt, flag_reject, 0, 1, 0, None, 0, 0, C.byref(st))
    assert rc == 0, L.oracle_last_error()
    return st


def oracle_counts(path, mapq_gt=0, flag_reject=0x600, min_bq=0, threads=1, window=None):
    """Closed-form per-read scatter oracle: counts[7, n] over the linear window (default: whole genome)."""
    L = oracle()
    _, tot, _, _ = oracle_info(path)
    a, b = window if window is not None else (0, tot)
    out = np.zeros((7, max(1, b - a)), np.uint32)
    st = ScatterStats()
    rc = L.oracle_base_counts(path.encode(), mapq_gt, flag_reject, min_bq, threads, 0, out.ctypes.data_as(C.c_void_p), a, max(1, b - a), C.byref(st))
    assert rc == 0, L.oracle_last_error()
    return out[:, :b - a], st


def oracle_segment_stats(path, seg_a, seg_b, thresholds=(), mapq_gt=0, flag_reject=0x600, min_bq=0, threads=1):
    """Closed-form region / window statistics for sorted, disjoint segments in linear coordinates:
    (n_reads[n], n_bases[n], cov_ge[n_thr, n])."""
    L = oracle()
    a = np.ascontiguousarray(seg_a, np.uint64)
    b = np.ascontiguousarray(seg_b, np.uint64)
    n = len(a)
    thr = np.ascontiguousarray(thresholds, np.uint32)
    reads, bases, cov = np.zeros(max(1, n), np.uint32), np.zeros(max(1, n), np.uint32), np.zeros((max(1, len(thr)), max(1, n)), np.uint32)
    rc = L.oracle_segment_stats(path.encode(), mapq_gt, flag_reject, min_bq, threads, n, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), len(thr),
                                thr.ctypes.data_as(C.c_void_p), reads.ctypes.data_as(C.c_void_p), bases.ctypes.data_as(C.c_void_p), cov.ctypes.data_as(C.c_void_p))
    assert rc == 0, L.oracle_last_error()
    return reads[:n], bases[:n], cov[:len(thr), :n]


def oracle_build_bai(path, threads=4):
    """The BAI index as the reference's IndexBuilder computes it (oracle_build_bai in oracle/depth_oracle.c), bins ascending."""
    import ctypes as C
    o = oracle()
    o.oracle_build_bai.restype = C.c_int64
    o.oracle_build_bai.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_int]
    n = o.oracle_build_bai(path.encode(), None, 0, threads)
    if n < 0:
        o.oracle_last_error.restype = C.c_char_p
        raise RuntimeError(o.oracle_last_error().decode())
    buf = (C.c_uint8 * n)()
    assert o.oracle_build_bai(path.encode(), buf, n, threads) == n
    return bytes(buf)


def parse_bai(b):
    """A .bai as ([(bins: {bin: [(beg, end)...]}, linear: [...])...], n_no_coor) -- independent of the order the bins were written in."""
    import struct
    assert b[:4] == b"BAI\x01"
    n_ref, = struct.unpack_from("<i", b, 4)
    o = 8
    refs = []
    for _ in range(n_ref):
        n_bin, = struct.unpack_from("<i", b, o)
        o += 4
        bins = {}
        for _ in range(n_bin):
            bin_, n_chunk = struct.unpack_from("<Ii", b, o)
            o += 8
            assert bin_ not in bins
            bins[bin_] = [struct.unpack_from("<QQ", b, o + 16 * k) for k in range(n_chunk)]
            o += 16 * n_chunk
        n_intv, = struct.unpack_from("<i", b, o)
        o += 4
        refs.append((bins, list(struct.unpack_from("<%dQ" % n_intv, b, o))))
        o += 8 * n_intv
    tail = None
    if o + 8 <= len(b):
        tail, = struct.unpack_from("<Q", b, o)
        o += 8
    assert o == len(b)
    return refs, tail


def oracle_counts_fix_mates(path, mapq_gt=0, flag_reject=0x600, min_bq=0, window=None):
    """Closed form for `-m` (base mode): counts[7, n] over the linear window and the number of (pair, column) fixes."""
    L = oracle()
    n_ref, total, ulen, nblk = oracle_info(path)
    a, b = (0, total) if window is None else window
    counts = np.zeros((7, max(0, b - a)), np.uint32)
    npc = C.c_uint64()
    rc = L.oracle_base_counts_fix_mates(path.encode(), mapq_gt, flag_reject, min_bq, counts.ctypes.data_as(C.c_void_p), a, b - a, C.byref(npc))
    assert rc == 0, L.oracle_last_error()
    return counts, npc.value


def interesting_window(path, pad=2000, **kw):
    """Linear window around the passing reads (keeps tests on human-sized headers small)."""
    st = oracle_scan(path, **kw)
    _, tot, _, _ = oracle_info(path)
    if st.n_pass == 0:
        return (0, min(tot, 4096))
    return (max(0, st.min_lin - pad), min(tot, st.max_lin + pad))


def oracle_cli(args, stdin=None):
    r = subprocess.run([ORACLE_EXE, "depth"] + list(args), capture_output=True)
    return r.returncode, r.stdout, r.stderr


def run_cli(args):
    r = subprocess.run([CLI] + list(args), capture_output=True)
    return r.returncode, r.stdout, r.stderr


def gen_bam(path, *extra):
    subprocess.check_call([BAMGEN, "-o", path] + [str(x) for x in extra], stderr=subprocess.DEVNULL)
    return path


def parse_records(u, first_off):
    """Pure-numpy/py walk of an inflated BAM stream -> list of (off, ref, pos, flag, mapq, n_cigar, span)."""
    import struct
    out = []
    o = first_off
    n = len(u)
    b = u.tobytes()
    while o + 4 <= n:
        bs = struct.unpack_from("<i", b, o)[0]
        if o + 4 + bs > n:
            break
        ref, pos, bmn, fnc, lseq = struct.unpack_from("<iiIIi", b, o + 4)
        lname = bmn & 0xFF
        ncig = fnc & 0xFFFF
        span = 0
        for k in range(ncig):
            c = struct.unpack_from("<I", b, o + 36 + lname + 4 * k)[0]
            if (c & 15) in (0, 2, 3, 7, 8):
                span += c >> 4
        out.append((o, ref, pos, fnc >> 16, (bmn >> 8) & 0xFF, ncig, span))
        o += 4 + bs
    return out


def header_first_record_offset(u):
    import struct
    b = u.tobytes()
    l_text = struct.unpack_from("<i", b, 4)[0]
    off = 8 + l_text
    n_ref = struct.unpack_from("<i", b, off)[0]
    off += 4
    refs = []
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", b, off)[0]
        name = b[off + 4:off + 4 + ln - 1].decode()
        L = struct.unpack_from("<i", b, off + 4 + ln)[0]
        refs.append((name, L))
        off += 8 + ln
    return off, refs


def reg2bin(beg, end):
    """The BAI bin of [beg, end) (SAM specification; BioD/bio/std/hts/bam/bai/bin.d:82-92)."""
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def write_bam(path, refs, reads, rg=None, block=0xFF00, level=6, quals=None, tags=None, bins=None, index=True):
    """Minimal BAM + dummy BAI writer for hand-made reads: (ref, pos, mapq, flag, cigar[(len,op)], seq, name).
    quals: optional list of per-read quality lists (default 30 everywhere); tags: optional per-read raw aux bytes;
    bins: optional per-read bin fields (default 4680 everywhere; "auto": reg2bin of the alignment); index=False: no .bai."""
    import struct
    import zlib
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in refs)
    if rg:
        text += "".join(f"@RG\tID:{i}\tSM:{s}\n" for i, s in rg)
    body = b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(refs))
    for n, l in refs:
        body += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l)
    code = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
    for ri, (ref, pos, mapq, flag, cigar, seq, name) in enumerate(reads):
        nm = name.encode() + b"\0"
        packed = bytearray()
        for i in range(0, len(seq), 2):
            packed.append((code[seq[i]] << 4) | (code[seq[i + 1]] if i + 1 < len(seq) else 0))
        if bins is None:
            bin_ = 4680
        elif bins == "auto":
            span = 0 if flag & 4 else sum(l for l, op in cigar if op in (0, 2, 3, 7, 8))
            bin_ = reg2bin(max(pos, 0), max(pos, 0) + max(span, 1))
        else:
            bin_ = bins[ri]
        rec = struct.pack("<iiIIiiii", ref, pos, (bin_ << 16) | (mapq << 8) | len(nm), (flag << 16) | len(cigar), len(seq), -1, -1, 0)
        rec += nm + b"".join(struct.pack("<I", (l << 4) | op) for l, op in cigar) + bytes(packed) + (bytes(quals[ri]) if quals else bytes([30] * len(seq))) + (tags[ri] if tags else b"")
        body += struct.pack("<i", len(rec)) + rec
    with open(path, "wb") as f:
        for i in range(0, len(body), block):
            chunk = body[i:i + block]
            c = zlib.compressobj(level, zlib.DEFLATED, -15)
            d = c.compress(chunk) + c.flush()
            f.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(d) + 25) + d + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
        f.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    if not index:
        return path
    with open(path + ".bai", "wb") as f:      # empty-but-valid index: depth only checks that it exists (depth.d:1166)
        f.write(b"BAI\1" + struct.pack("<i", len(refs)) + b"".join(struct.pack("<ii", 0, 0) for _ in refs) + struct.pack("<Q", 0))
    return path




def write_bgzf(path, body, n_ref, block=0xFF00, level=6):
    """BGZF-compress an inflated BAM stream and put an empty-but-valid index next to it."""
    import struct
    import zlib
    with open(path, "wb") as f:
        for i in range(0, len(body), block):
            chunk = body[i:i + block]
            c = zlib.compressobj(level, zlib.DEFLATED, -15)
            d = c.compress(chunk) + c.flush()
            f.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(d) + 25) + d + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
        f.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    with open(path + ".bai", "wb") as f:
        f.write(b"BAI\1" + struct.pack("<i", n_ref) + b"".join(struct.pack("<ii", 0, 0) for _ in range(n_ref)) + struct.pack("<Q", 0))
    return path


def subset_bam(src, dst, keep):
    """Copy of src holding only the records i (file order) with keep[i] true -- what a read filter leaves."""
    import struct
    u = oracle_inflate(src)
    first, refs = header_first_record_offset(u)
    b = u.tobytes()
    out = [b[:first]]
    o, i = first, 0
    while o + 4 <= len(b):
        bs = struct.unpack_from("<i", b, o)[0]
        if o + 4 + bs > len(b):
            break
        if keep[i]:
            out.append(b[o:o + 4 + bs])
        o += 4 + bs
        i += 1
    assert i == len(keep)
    return write_bgzf(dst, b"".join(out), len(refs))
