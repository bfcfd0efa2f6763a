"""CPU-side check of the device DEFLATE logic (sambamba_b200/csrc/inflate_core.cuh compiled for the
host, one emulated lane) against zlib -- the library the reference itself calls
(BioD/bio/core/bgzf/block.d:162-183).  No GPU needed."""
import ctypes as C
import os
import random
import zlib

import numpy as np
import pytest

import helpers
from helpers import GOLDEN, ROOT


class TwoPhase:
    """The two-phase decoder (inflate2_core.cuh: phase 1 as one lane, phase 2 in its serial restatement) behind the interface of
    the one-phase harness; a block phase 1 hands back (INF_FALLBACK = 100) goes to the one-phase decoder, as in the pipeline."""

    def __init__(self, L, lims):
        self.L, self.lims = L, lims
        self.fallbacks = 0

    def emul_inflate_file(self, path, dst, cap, fe):
        self.L.emul_set_lims(self.lims)
        st = (C.c_ulonglong * 4)()
        n = self.L.emul_inflate2_file(path, dst, cap, fe, st, C.c_uint32(4))
        self.fallbacks += st[2]
        return n

    def emul_inflate_raw(self, s, n, dst, isize, off, ooff):
        self.L.emul_set_lims(self.lims)
        rc = self.L.emul_inflate2_raw(s, n, dst, isize, off, ooff)
        if rc == 100:
            self.fallbacks += 1
            rc = self.L.emul_inflate_raw(s, n, dst, isize, off, ooff)
        return rc


@pytest.fixture(scope="module", params=[0, 1, 2, 3], ids=["one-phase", "one-phase-three-literals", "two-phase", "two-phase-limits-in-table-storage"])
def em(request):
    """The lane logic of every K1 kernel: the one-phase decoder (k1_inflate: header blocks, k1_fallback) and the two-phase one (k1_huff + k1_lz)."""
    L = C.CDLL(os.path.join(ROOT, "tests", "emul", "libemul.so"))
    L.emul_inflate_file.restype = C.c_long
    L.emul_inflate2_file.restype = C.c_long
    L.emul_inflate_raw.restype = C.c_int
    L.emul_inflate2_raw.restype = C.c_int
    if request.param >= 2:
        yield TwoPhase(L, request.param - 2)
        return
    L.emul_set_lit3(request.param)
    yield L
    L.emul_set_lit3(0)


@pytest.fixture(scope="module")
def synth(tmp_path_factory):
    d = tmp_path_factory.mktemp("emul")
    return [helpers.gen_bam(str(d / "tiny.bam"), "--preset", "tiny", "-t", 2),
            helpers.gen_bam(str(d / "mix.bam"), "--preset", "tiny", "-n", 12000, "--stored-every", 5, "-t", 2)]


def test_bam_files_match_zlib(em, synth):
    files = [os.path.join(GOLDEN, f) for f in os.listdir(GOLDEN) if f.endswith(".bam")] + synth
    for p in files:
        want = helpers.oracle_inflate(p)
        got = np.zeros(len(want) + 64, np.uint8)
        fe = C.c_int()
        n = em.emul_inflate_file(p.encode(), got.ctypes.data_as(C.c_void_p), C.c_uint64(len(want)), C.byref(fe))
        assert n == len(want) and fe.value == 0, (p, n, fe.value)
        assert np.array_equal(got[:n], want), p
        assert not got[n:].any(), "wrote past the end"


def _cases():
    rnd = random.Random(1)
    cases = {
        "empty": b"", "one": b"a", "zeros": bytes(65280), "text": (b"the quick brown fox jumps over the lazy dog " * 1500)[:65280],
        "rand": bytes(rnd.getrandbits(8) for _ in range(65280)),
        "skew": bytes(min(255, int(rnd.expovariate(0.02))) for _ in range(65280)),
        "allsyms": bytes(range(256)) * 200,
        "runs": b"".join(bytes([rnd.getrandbits(8)]) * rnd.randint(1, 300) for _ in range(400))[:65280],
        "period3": (b"abc" * 22000)[:65280], "period7": (b"abcdefg" * 9500)[:65280], "period9": (b"abcdefghi" * 7300)[:65280],
    }
    geo = bytearray()
    for _ in range(60000):
        k = 0
        while rnd.random() < 0.5 and k < 255:
            k += 1
        geo.append((k * 37) & 255)
    cases["geometric"] = bytes(geo)
    return cases


def test_crafted_streams(em):
    strategies = [(6, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED),
                  (6, zlib.Z_HUFFMAN_ONLY), (0, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_RLE)]
    for name, d in _cases().items():
        for lvl, strat in strategies:
            c = zlib.compressobj(lvl, zlib.DEFLATED, -15, 8, strat)
            s = c.compress(d) + c.flush()
            for off, ooff in ((0, 0), (1, 5), (2, 16), (3, 31)):
                out = np.zeros(len(d) + ooff + 32, np.uint8)
                rc = em.emul_inflate_raw(s, len(s), out.ctypes.data_as(C.c_void_p), len(d), off, ooff)
                assert rc == 0, (name, lvl, strat, off, rc)
                assert bytes(out[ooff:ooff + len(d)]) == d, (name, lvl, strat, off)
                assert not out[:ooff].any() and not out[ooff + len(d):].any(), "wrote outside its block"


def test_sync_flush_and_multi_block(em):
    # Z_SYNC_FLUSH inserts empty stored blocks; Z_FULL_FLUSH resets the window
    rnd = random.Random(3)
    d = bytes(rnd.choice(b"ACGT") for _ in range(30000))
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    s = c.compress(d[:10000]) + c.flush(zlib.Z_SYNC_FLUSH) + c.compress(d[10000:20000]) + c.flush(zlib.Z_FULL_FLUSH) + c.compress(d[20000:]) + c.flush()
    out = np.zeros(len(d), np.uint8)
    assert em.emul_inflate_raw(s, len(s), out.ctypes.data_as(C.c_void_p), len(d), 0, 0) == 0
    assert bytes(out) == d


def test_corrupt_streams_fail_cleanly(em):
    d = (b"hello world, hello world! " * 500)
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    s = bytearray(c.compress(d) + c.flush())
    out = np.zeros(len(d) + 64, np.uint8)
    # wrong isize
    assert em.emul_inflate_raw(bytes(s), len(s), out.ctypes.data_as(C.c_void_p), len(d) - 1, 0, 0) != 0
    assert em.emul_inflate_raw(bytes(s), len(s), out.ctypes.data_as(C.c_void_p), len(d) + 1, 0, 0) != 0
    # bit flips must never crash or write out of bounds; they either fail or produce len(d) bytes
    rnd = random.Random(5)
    for _ in range(300):
        t = bytearray(s)
        i = rnd.randrange(len(t))
        t[i] ^= 1 << rnd.randrange(8)
        out[:] = 0
        em.emul_inflate_raw(bytes(t), len(t), out.ctypes.data_as(C.c_void_p), len(d), 0, 0)
        assert not out[len(d):].any()
    # reserved block type
    assert em.emul_inflate_raw(bytes([0x07]), 1, out.ctypes.data_as(C.c_void_p), 0, 0, 0) != 0


def test_two_phase_hands_back_what_it_cannot_express():
    """More deflate blocks in one BGZF block than phase 2 keeps tables for (MAX_SEG = 8), or more matches than the token area
    holds: phase 1 answers INF_FALLBACK (100) and the one-phase decoder takes the block (k1_fallback in the pipeline)."""
    L = C.CDLL(os.path.join(ROOT, "tests", "emul", "libemul.so"))
    L.emul_inflate2_raw.restype = C.c_int
    L.emul_inflate_raw.restype = C.c_int
    L.emul_inflate2_file.restype = C.c_long
    rnd = random.Random(11)
    d = bytes(rnd.choice(b"ACGTN") for _ in range(24000))
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    s = b"".join(c.compress(d[i:i + 2000]) + c.flush(zlib.Z_FULL_FLUSH) for i in range(0, 24000, 2000)) + c.flush()      # 12 dynamic blocks + empty stored ones
    out = np.zeros(len(d), np.uint8)
    assert L.emul_inflate2_raw(s, len(s), out.ctypes.data_as(C.c_void_p), len(d), 0, 0) == 100
    assert L.emul_inflate_raw(s, len(s), out.ctypes.data_as(C.c_void_p), len(d), 0, 0) == 0 and bytes(out) == d
    # eight segments are still fine
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    s = b"".join(c.compress(d[i:i + 3000]) + c.flush(zlib.Z_FULL_FLUSH) for i in range(0, 9000, 3000)) + c.compress(d[9000:]) + c.flush()
    out[:] = 0
    assert L.emul_inflate2_raw(s, len(s), out.ctypes.data_as(C.c_void_p), len(d), 0, 0) == 0 and bytes(out) == d
    # token area: a BGZF file whose blocks are nothing but 3-byte matches at changing distances needs isize / 3 tokens, more than the isize / 4 + 16 the pipeline provides
    import struct
    body = bytearray(rnd.getrandbits(8) for _ in range(300))
    while len(body) < 60000:
        k = rnd.randrange(0, len(body) - 3)
        body += body[k:k + 3]
    path = "/tmp/bdepth_tok_overflow_%d.bgzf" % os.getpid()
    with open(path, "wb") as f:
        cz = zlib.compressobj(9, zlib.DEFLATED, -15)
        z = cz.compress(bytes(body)) + cz.flush()
        f.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(z) + 25) + z + struct.pack("<II", zlib.crc32(bytes(body)), len(body)))
        f.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    try:
        got = np.zeros(len(body) + 64, np.uint8)
        fe = C.c_int()
        st = (C.c_ulonglong * 4)()
        n = L.emul_inflate2_file(path.encode(), got.ctypes.data_as(C.c_void_p), C.c_uint64(len(body)), C.byref(fe), st, C.c_uint32(4))
        if st[0] > len(body) // 4 + 16 or n < 0:       # zlib found that many matches: phase 1 must have refused, not overrun
            assert fe.value == 100 and st[2] == 1
        else:
            assert n == len(body) and bytes(got[:n]) == bytes(body)
    finally:
        os.unlink(path)


def test_phase_two_warp_code_under_the_emulation(tmp_path):
    """k1_lz as the GPU runs it -- warp scans, literal scatter, pointer jumping -- on every fixture and on inputs with stored
    blocks, through the emulated library with BDEPTH_EMU_K1LZ_WARP=1 (the big emulated suites use the serial restatement of
    phase 2: a warp collective costs 32 fiber switches there)."""
    import subprocess
    import sys
    mix = helpers.gen_bam(str(tmp_path / "mix.bam"), "--preset", "tiny", "-n", 6000, "--stored-every", 4, "-t", 2)
    # long self-overlapping matches (distance 1..3, length 258), period-sized distances, identical records: what zlib makes of constant data
    rep = [(0, 10 + (i // 7), 60, 0, [(120, 0)], "A" * 120, "q") for i in range(900)] + [(0, 400 + i, 60, 0, [(60, 0)], "ACG" * 20, "n%d" % (i % 3)) for i in range(600)]
    runs = helpers.write_bam(str(tmp_path / "runs.bam"), [("c", 5000)], rep, quals=[[30] * len(r[5]) for r in rep], level=9)
    code = (
        "import sys, os, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import helpers, sambamba_b200._lib as L\n"
        "L.lib_path = lambda: %r\n"
        "import sambamba_b200 as sb\n"
        "for p in sys.argv[1:]:\n"
        "    with sb.BDepth(p) as b:\n"
        "        assert np.array_equal(b.inflate(), helpers.oracle_inflate(p)), p\n"
        "print('ok')\n") % (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emul", "libbdepth_emul.so"))
    files = [os.path.join(GOLDEN, f) for f in sorted(os.listdir(GOLDEN)) if f.endswith(".bam")] + [mix, runs]
    for variant in ({}, {"BDEPTH_K1LZ": "v12"}, {"BDEPTH_K1LZ": "flat"}):        # k1_lz (lane = token, dependency rounds; compacted and full literal table) and k1_lz_flat (lane = output byte, pointer jumping)
        r = subprocess.run([sys.executable, "-c", code] + files, env=dict(os.environ, BDEPTH_EMU_K1LZ_WARP="1", **variant), capture_output=True, text=True, timeout=900)
        assert r.returncode == 0 and r.stdout.strip() == "ok", (variant, r.stderr[-2000:])
