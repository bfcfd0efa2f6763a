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


@pytest.fixture(scope="module", params=[0, 1], ids=["two-literals", "three-literals"])
def em(request):
    """Both variants of the lane logic: the default one and the one behind k1_inflate_lit3 (BDEPTH_K1_LIT3=1)."""
    L = C.CDLL(os.path.join(ROOT, "tests", "emul", "libemul.so"))
    L.emul_inflate_file.restype = C.c_long
    L.emul_inflate_raw.restype = C.c_int
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
