"""Size-independent properties of the hot path, used by tests/test_zz_gpu_fullsize.py at BASELINE configs[1] size and
checked on the CPU at small size (tests/test_fullsize_props_cpu.py) with the oracle's outputs standing in for the GPU's.
Each takes plain arrays, so the same code judges both."""
import struct
import zlib

import numpy as np


def bgzf_blocks(raw):
    """(offset, block size, crc32, isize) of every BGZF member of a file image (bytes-like)."""
    out, off, n = [], 0, len(raw)
    while off + 18 <= n:
        xlen = raw[off + 10] | (raw[off + 11] << 8)
        bsize, l = None, 0
        while l < xlen:
            slen = raw[off + 14 + l] | (raw[off + 15 + l] << 8)
            if raw[off + 12 + l] == 66 and raw[off + 13 + l] == 67:
                bsize = (raw[off + 16 + l] | (raw[off + 17 + l] << 8)) + 1
            l += 4 + slen
        assert bsize, "not a BGZF member"
        crc, isize = struct.unpack_from("<II", raw, off + bsize - 8)
        out.append((off, bsize, crc, isize))
        off += bsize
    return out


def inflate_matches_the_files_own_checksums(raw, u):
    """Checksum of checksums: the CRC32 each BGZF member stores for its payload (the reference never checks it in release
    builds) against the CRC32 of the inflated bytes."""
    pos, bad = 0, 0
    mv = memoryview(u)
    for _, _, crc, isize in bgzf_blocks(raw):
        if zlib.crc32(mv[pos:pos + isize]) != crc:
            bad += 1
        pos += isize
    return pos == len(u) and bad == 0


def scan_is_sorted_and_consistent(cols, n_ref):
    """Coordinate order of the scanned records and sanity of the decoded columns."""
    ref = cols["ref_id"].astype(np.int64)
    pos = cols["pos"].astype(np.int64)
    key = np.where(ref < 0, np.int64(n_ref), ref) * (1 << 32) + np.where(pos < 0, 0, pos)
    return bool(np.all(key[1:] >= key[:-1])) and bool(np.all(ref < n_ref)) and bool(np.all(np.diff(cols["rec_off"].astype(np.int64)) >= 36))


def passing(cols, mapq_gt=0, flag_reject=0x600):
    f = cols["flag"].astype(np.uint32)
    return (cols["ref_id"] >= 0) & (cols["pos"] >= 0) & (cols["mapq"].astype(np.int32) > mapq_gt) & ((f & flag_reject) == 0) & ((f & 4) == 0) & (cols["span"] > 0)


def counters_add_up(cols, counts, covered_positions):
    """Without -q every reference base a passing read covers lands in exactly one of the 7 planes: the grand total equals
    the sum of the reads' reference spans (clipped to the reference); the run's covered-position count equals the
    number of non-zero columns."""
    ok = passing(cols)
    total_span = int(cols["span"][ok].astype(np.uint64).sum())
    col = counts.sum(axis=0, dtype=np.uint64)
    return int(col.sum()) == total_span and int(np.count_nonzero(col)) == int(covered_positions)
