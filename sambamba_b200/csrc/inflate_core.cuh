// inflate_core.cuh -- raw DEFLATE (RFC 1951) decoder for one BGZF block, written so that ONE
// CUDA LANE decodes ONE block with its Huffman tables in (lane-interleaved) shared memory.
//
// Replaces: BioD/bio/core/bgzf/block.d:127-216 (decompressBgzfBlock -> zlib inflateInit2(-15)/
// inflate(Z_FINISH)/inflateEnd, bound through BioD/bio/core/utils/zlib.d:143-162).  The
// arithmetic being restated is zlib's (third-party, not under /root/reference); RFC 1951 output
// is deterministic so any correct inflater is bit-identical.
//
// Design (see DESIGN.md "K1").  A warp-per-block decoder issues ~150 instructions per symbol
// with one useful lane; giving every lane its own block makes those instructions decode 32
// symbols.  Consequences:
//  * Tables must be tiny (they bound occupancy), so there is NO lookup table: decoding is
//    canonical-Huffman arithmetic.  Measured on BAM data a first-level LUT misses on 5-10 % of
//    symbols, i.e. in ~97 % of 32-lane iterations, so its fast path would almost never save the
//    warp anything while costing 4x the shared memory.
//      len = 1 + #{ j in 1..14 : peek15 >= lim[j] }      lim[] (left-justified) in REGISTERS
//      sym = sorted[ (peek15 >> (15-len)) + delta[len] ]  two shared-memory loads
//  * The SIMT loop is a flat state machine: each iteration a lane either decodes ONE symbol or
//    moves ONE 8-byte round of a pending LZ77 copy, so no lane waits for another lane's match.
//  * Output goes through a 256-byte per-lane ring in shared memory: literals are byte stores to
//    shared memory, completed 16-byte lines leave as one aligned 16-byte global store, and
//    matches with distance <= 240 are served from the ring.  Far matches read global memory; the
//    loaded bytes are stored one iteration later, after the next symbol has been decoded, so the
//    L2 round trip overlaps the Huffman arithmetic.
//  per lane: 288 x u16 litlen symbols (144 w) + 15 x i16 delta (8 w) + 32 x u8 dist symbols (8 w)
//            + delta (8 w) + ring (64 w) = 232 words = 928 B; 29 KB per warp; 7 warps per SM.
//
// The code is __host__ __device__ so the exact same logic is unit-tested on the CPU against zlib
// (tests/test_emul_inflate.py); the product only ever calls it from kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BD_HD __host__ __device__ __forceinline__
#else
#define BD_HD inline
#endif

namespace bdk {

#ifdef BD_INFLATE_STATS
struct InflateStats { unsigned long long lits, matches, match_bytes, tables, near_matches, iters, len_hist[16], mlen_hist[16]; };
extern InflateStats g_inflate_stats;
#define BD_STAT(x) (x)
#else
#define BD_STAT(x) ((void)0)
#endif

enum InflateStatus : int {
    INF_OK = 0,
    INF_ERR_BTYPE = 2,       // reserved block type 3
    INF_ERR_STORED = 3,      // LEN != ~NLEN
    INF_ERR_TABLE = 4,       // over-subscribed / incomplete / missing EOB / bad repeat
    INF_ERR_CODE = 5,        // invalid code in stream
    INF_ERR_DIST = 6,        // distance too far back
    INF_ERR_OVERRUN = 7,     // output would exceed isize
    INF_ERR_SHORT = 8,       // stream ended with fewer than isize bytes
    INF_ERR_INPUT = 9        // ran past the end of the compressed data
};

// word offsets of the per-lane table regions
constexpr int T_LL_SYMS = 0;       // 288 x u16
constexpr int T_LL_DELTA = 144;    // 15 x i16 (index len-1), padded to 16
constexpr int T_D_SYMS = 152;      // 32 x u8
constexpr int T_D_DELTA = 160;     // 15 x i16
constexpr int T_RING = 168;        // 256-byte output ring
constexpr int RING_WORDS = 64;
constexpr int T_WORDS = T_RING + RING_WORDS;   // 232
constexpr uint32_t NEAR_MAX = 240;             // matches at most this far back are served from the ring

// ---- table storage policies -------------------------------------------------------------
// Lane-interleaved shared memory: word w of this lane lives at base[w * 32]; every lane always
// hits its own bank, so data-dependent indices never conflict.
struct SmemTab {
    uint32_t* base;
    BD_HD uint32_t ldw(int w) const { return base[w * 32]; }
    BD_HD void stw(int w, uint32_t v) const { base[w * 32] = v; }
    BD_HD void ring_st8(uint32_t a, uint32_t v) const {
        reinterpret_cast<uint8_t*>(base + T_RING * 32)[(((a >> 2) & (RING_WORDS - 1)) << 7) | (a & 3)] = (uint8_t)v;
    }
    BD_HD uint32_t ring_ld8(uint32_t a) const {
        return reinterpret_cast<const uint8_t*>(base + T_RING * 32)[(((a >> 2) & (RING_WORDS - 1)) << 7) | (a & 3)];
    }
    BD_HD uint32_t ring_ldw(uint32_t widx) const { return base[(T_RING + (widx & (RING_WORDS - 1))) * 32]; }
};
// Plain array (host tests).
struct FlatTab {
    uint32_t* base;
    BD_HD uint32_t ldw(int w) const { return base[w]; }
    BD_HD void stw(int w, uint32_t v) const { base[w] = v; }
    BD_HD void ring_st8(uint32_t a, uint32_t v) const { reinterpret_cast<uint8_t*>(base + T_RING)[a & (RING_WORDS * 4 - 1)] = (uint8_t)v; }
    BD_HD uint32_t ring_ld8(uint32_t a) const { return reinterpret_cast<const uint8_t*>(base + T_RING)[a & (RING_WORDS * 4 - 1)]; }
    BD_HD uint32_t ring_ldw(uint32_t widx) const { return base[T_RING + (widx & (RING_WORDS - 1))]; }
};

template <class Tab> BD_HD uint32_t tab_ld16(const Tab& t, int region, int i) {
    uint32_t w = t.ldw(region + (i >> 1));
    return (i & 1) ? (w >> 16) : (w & 0xFFFFu);
}
template <class Tab> BD_HD void tab_st16(const Tab& t, int region, int i, uint32_t v) {
    int wi = region + (i >> 1);
    uint32_t w = t.ldw(wi);
    w = (i & 1) ? ((w & 0x0000FFFFu) | (v << 16)) : ((w & 0xFFFF0000u) | (v & 0xFFFFu));
    t.stw(wi, w);
}
template <class Tab> BD_HD uint32_t tab_ld8(const Tab& t, int region, int i) {
    return (t.ldw(region + (i >> 2)) >> ((i & 3) * 8)) & 0xFFu;
}
template <class Tab> BD_HD void tab_st8(const Tab& t, int region, int i, uint32_t v) {
    int wi = region + (i >> 2), sh = (i & 3) * 8;
    uint32_t w = t.ldw(wi);
    t.stw(wi, (w & ~(0xFFu << sh)) | ((v & 0xFFu) << sh));
}

BD_HD uint32_t bitrev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}

// ---- bit reader: 64-bit reservoir fed by aligned 32-bit words, next word always prefetched --
struct BitReader {
    const uint32_t* wp;     // address of the prefetched word
    const uint32_t* wend;   // one past the last word that may contain stream bits
    uint64_t bb;
    uint32_t nw;            // prefetched next word
    int bc;
    BD_HD void init(const uint32_t* words, uint64_t byte_off, uint32_t nbytes) {
        wp = words + (byte_off >> 2);
        wend = words + ((byte_off + nbytes + 3) >> 2);
        unsigned mis = (unsigned)(byte_off & 3);
        bb = (uint64_t)(*wp++) >> (8 * mis);
        bc = 32 - 8 * (int)mis;
        nw = *wp;
    }
    BD_HD void refill() {   // afterwards bc >= 33
        if (bc <= 32) { bb |= (uint64_t)nw << bc; bc += 32; nw = *++wp; }
    }
    BD_HD uint32_t peek(int n) const { return (uint32_t)bb & ((1u << n) - 1u); }
    BD_HD void drop(int n) { bb >>= n; bc -= n; }
    BD_HD uint32_t get(int n) { uint32_t v = peek(n); drop(n); return v; }
    // the reservoir + prefetch run at most three words past the last stream word
    BD_HD bool overrun() const { return wp > wend + 3; }
};

// Left-justified 15-bit limits, index len-1.  v[14] == 32768 for a complete code.
struct HuffLim { uint32_t v[15]; };

// Build one canonical table from code lengths lens[0..n).  KIND 0 = litlen (u16 symbols),
// 1 = dist (u8 symbols).  Mirrors zlib inflate_table()'s validity rules: over-subscribed ->
// error; incomplete -> error unless the longest code is <= 1 bit (dist may also be empty).
template <class Tab, int KIND>
BD_HD int build_table(const Tab& t, const uint8_t* lens, int n, HuffLim& lim) {
    uint32_t cnt[16];
#pragma unroll
    for (int i = 0; i < 16; i++) cnt[i] = 0;
    for (int s = 0; s < n; s++) cnt[lens[s]]++;
    int left = 1, maxl = 0;
#pragma unroll
    for (int l = 1; l <= 15; l++) { left <<= 1; left -= (int)cnt[l]; if (cnt[l]) maxl = l; }
    {   // over-subscription can only be detected reliably per level; redo exactly
        int lf = 1; bool over = false;
#pragma unroll
        for (int l = 1; l <= 15; l++) { lf <<= 1; lf -= (int)cnt[l]; if (lf < 0) over = true; }
        if (over) return INF_ERR_TABLE;
    }
    if (left > 0 && (KIND == 0 ? maxl != 1 : maxl > 1)) return INF_ERR_TABLE;
    uint32_t nxt[16];      // next free slot in the sorted table per length
    uint32_t code = 0, o = 0;
#pragma unroll
    for (int l = 1; l <= 15; l++) {
        nxt[l] = o;
        lim.v[l - 1] = (code + cnt[l]) << (15 - l);
        int delta = (int)o - (int)code;
        tab_st16(t, KIND == 0 ? T_LL_DELTA : T_D_DELTA, l - 1, (uint32_t)delta & 0xFFFFu);
        o += cnt[l];
        code = (code + cnt[l]) << 1;
    }
    for (int s = 0; s < n; s++) {
        uint32_t l = lens[s];
        if (!l) continue;
        if (KIND == 0) tab_st16(t, T_LL_SYMS, (int)nxt[l], (uint32_t)s); else tab_st8(t, T_D_SYMS, (int)nxt[l], (uint32_t)s);
        nxt[l]++;
    }
    return INF_OK;
}

// Decode one symbol of a canonical code.  Returns -1 on an invalid code; consumes its bits.
template <class Tab, int KIND>
BD_HD int decode_sym(const Tab& t, const HuffLim& lim, BitReader& br) {
    uint32_t rev15 = bitrev32((uint32_t)br.bb) >> 17;
    // 14 independent compares summed as a tree (a serial += chain would cost 14 dependent adds)
#define BD_GE(j) ((rev15 >= lim.v[j]) ? 1 : 0)
    int s0 = (BD_GE(0) + BD_GE(1)) + (BD_GE(2) + BD_GE(3));
    int s1 = (BD_GE(4) + BD_GE(5)) + (BD_GE(6) + BD_GE(7));
    int s2 = (BD_GE(8) + BD_GE(9)) + (BD_GE(10) + BD_GE(11));
    int s3 = (BD_GE(12) + BD_GE(13)) + 1;
#undef BD_GE
    int L = (s0 + s1) + (s2 + s3);
    if (rev15 >= lim.v[14]) return -1;
    int delta = (int)(int16_t)tab_ld16(t, KIND == 0 ? T_LL_DELTA : T_D_DELTA, L - 1);
    int idx = (int)(rev15 >> (15 - L)) + delta;
    if (idx < 0 || idx >= (KIND == 0 ? 288 : 32)) return -1;
    BD_STAT(KIND == 0 ? g_inflate_stats.len_hist[L]++ : 0);
    br.drop(L);
    return KIND == 0 ? (int)tab_ld16(t, T_LL_SYMS, idx) : (int)tab_ld8(t, T_D_SYMS, idx);
}

// Read the dynamic-block header and produce lens[] (RFC 1951 3.2.7).  lens must hold 320 bytes.
// The 19-symbol code-length code lives entirely in registers.
BD_HD int read_dynamic_lens(BitReader& br, uint8_t* lens, int& nlen, int& ndist) {
    br.refill();
    nlen = (int)br.get(5) + 257;
    ndist = (int)br.get(5) + 1;
    int ncode = (int)br.get(4) + 4;
    if (nlen > 286 || ndist > 30) return INF_ERR_TABLE;
    uint64_t cl = 0;    // 19 x 3-bit lengths indexed by symbol
    const uint8_t ord[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    for (int i = 0; i < ncode; i++) { br.refill(); cl |= (uint64_t)br.get(3) << (3 * ord[i]); }
    uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int s = 0; s < 19; s++) cnt[(cl >> (3 * s)) & 7]++;
    {
        int left = 1;
        for (int l = 1; l <= 7; l++) { left <<= 1; left -= (int)cnt[l]; if (left < 0) return INF_ERR_TABLE; }
        if (left > 0) return INF_ERR_TABLE;     // zlib: an incomplete code-length code is always an error
    }
    uint32_t first[8], lim[8], offs[8];
    { uint32_t code = 0, o = 0; for (int l = 1; l <= 7; l++) { first[l] = code; offs[l] = o; o += cnt[l]; lim[l] = (code + cnt[l]) << (7 - l); code = (code + cnt[l]) << 1; } }
    uint64_t sorted_lo = 0, sorted_hi = 0;   // 19 x 5-bit symbols sorted by (len, sym)
    {
        uint32_t nx[8]; for (int l = 1; l <= 7; l++) nx[l] = offs[l];
        for (int s = 0; s < 19; s++) {
            uint32_t l = (uint32_t)(cl >> (3 * s)) & 7; if (!l) continue;
            uint32_t pos = nx[l]++;
            if (pos < 12) sorted_lo |= (uint64_t)s << (5 * pos); else sorted_hi |= (uint64_t)s << (5 * (pos - 12));
        }
    }
    int n = 0, total = nlen + ndist;
    while (n < total) {
        br.refill();
        uint32_t rev7 = bitrev32((uint32_t)br.bb) >> 25;
        int L = 1;
#pragma unroll
        for (int j = 1; j <= 6; j++) L += (rev7 >= lim[j]) ? 1 : 0;
        if (rev7 >= lim[7]) return INF_ERR_TABLE;
        uint32_t idx = (rev7 >> (7 - L)) - first[L] + offs[L];
        if (idx >= 19) return INF_ERR_TABLE;
        uint32_t sym = (idx < 12) ? (uint32_t)(sorted_lo >> (5 * idx)) & 31u : (uint32_t)(sorted_hi >> (5 * (idx - 12))) & 31u;
        br.drop(L);
        if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
        uint32_t rep, val = 0;
        if (sym == 16) { if (n == 0) return INF_ERR_TABLE; val = lens[n - 1]; rep = 3 + br.get(2); }
        else if (sym == 17) rep = 3 + br.get(3);
        else rep = 11 + br.get(7);
        if (n + (int)rep > total) return INF_ERR_TABLE;
        while (rep--) lens[n++] = (uint8_t)val;
    }
    if (lens[256] == 0) return INF_ERR_TABLE;   // zlib: "invalid code -- missing end-of-block"
    return INF_OK;
}

// Output policy: the whole inflated stream as a byte array (global memory on the device).
// Addresses are ABSOLUTE stream offsets; put16 needs a 16-byte-aligned offset.
struct ByteOut {
    uint8_t* p;
    BD_HD void put(uint64_t i, uint32_t v) const { p[i] = (uint8_t)v; }
    BD_HD uint32_t get(uint64_t i) const { return p[i]; }
    BD_HD void put16(uint64_t i, uint32_t a, uint32_t b, uint32_t c, uint32_t d) const {
#if defined(__CUDA_ARCH__)
        *reinterpret_cast<uint4*>(p + i) = make_uint4(a, b, c, d);
#else
        uint32_t w[4] = {a, b, c, d};
        for (int k = 0; k < 16; k++) p[i + k] = (uint8_t)(w[k >> 2] >> ((k & 3) * 8));
#endif
    }
};

enum { ST_HDR = 0, ST_SYM = 1, ST_STORED = 2, ST_DONE = 3 };

// Warp convergence.  Lanes take different paths through an iteration (literal / match / header)
// and finish their blocks at different times.  With independent thread scheduling nothing forces
// them back together, and the first version of this kernel ran with 1.02 active threads per
// instruction (ncu, profiles/k1_v1_summary.md).  Every iteration therefore ends in a ballot over
// the lanes still decoding: it is both the reconvergence point and the next iteration's mask.
#if defined(__CUDA_ARCH__)
#define BD_BALLOT(mask, pred) __ballot_sync((mask), (pred))
#else
#define BD_BALLOT(mask, pred) ((pred) ? 1u : 0u)
#endif

// A 16-byte line [line, line+16) of the output has just been completed: send it to memory.
// Lines that begin before this block's first byte a0 are shared with the previous block, so only
// this block's bytes are written, one by one.
template <class Tab, class Out>
BD_HD void flush_line(const Tab& t, const Out& out, uint64_t line, uint64_t a0) {
    if (line >= a0) {
        uint32_t w = (uint32_t)(line >> 2);
        out.put16(line, t.ring_ldw(w), t.ring_ldw(w + 1), t.ring_ldw(w + 2), t.ring_ldw(w + 3));
    } else {
        for (uint64_t a = a0; a < line + 16; a++) out.put(a, t.ring_ld8((uint32_t)a));
    }
}

// Parse block headers until a block with content (or the end) is reached.  Returns an error code.
template <class Tab>
BD_HD int next_block(const Tab& t, BitReader& br, HuffLim& ll, HuffLim& dd, uint8_t* lens, uint32_t pos, uint32_t isize,
                     int& state, uint32_t& bfinal, uint32_t& stored_rem) {
    while (state == ST_HDR) {
        br.refill();
        bfinal = br.get(1); uint32_t btype = br.get(2);
        if (btype == 0) {
            br.drop(br.bc & 7);
            br.refill();
            uint32_t len = br.get(16); br.refill(); uint32_t nlen = br.get(16);
            if ((len ^ 0xFFFFu) != nlen) return INF_ERR_STORED;
            if (pos + len > isize) return INF_ERR_OVERRUN;
            stored_rem = len;
            state = len ? ST_STORED : (bfinal ? ST_DONE : ST_HDR);
        } else if (btype == 3) {
            return INF_ERR_BTYPE;
        } else {
            int nl, nd;
            if (btype == 1) {
                // fixed code (RFC 1951 3.2.6): complete over 288 litlen / 32 distance symbols; litlen 286/287
                // and distance 30/31 are rejected after decoding, as zlib does
                nl = 288; nd = 32;
                for (int i = 0; i < 144; i++) lens[i] = 8;
                for (int i = 144; i < 256; i++) lens[i] = 9;
                for (int i = 256; i < 280; i++) lens[i] = 7;
                for (int i = 280; i < 288; i++) lens[i] = 8;
                for (int i = 0; i < 32; i++) lens[288 + i] = 5;
            } else {
                int rc = read_dynamic_lens(br, lens, nl, nd);
                if (rc) return rc;
            }
            BD_STAT(g_inflate_stats.tables++);
            int rc = build_table<Tab, 0>(t, lens, nl, ll);
            if (rc) return rc;
            rc = build_table<Tab, 1>(t, lens + nl, nd, dd);
            if (rc) return rc;
            state = ST_SYM;
        }
        if (br.overrun()) return INF_ERR_INPUT;
    }
    return INF_OK;
}

// Inflate one raw-deflate stream.  words/byte_off/nbytes locate the compressed data inside a
// 4-byte-aligned buffer that has >= 64 readable bytes after the last block.  The block's isize
// output bytes go to absolute stream offsets [a0, a0+isize) of `out`.  On the device ALL 32 lanes
// of the warp must call this together; lanes without a block pass active = false.
template <class Tab, class Out>
BD_HD int inflate_block(const Tab& t, const uint32_t* words, uint64_t byte_off, uint32_t nbytes, const Out& out, uint64_t a0, uint32_t isize, uint8_t* lens /* 320 B scratch */, bool active = true) {
    unsigned mask = BD_BALLOT(0xFFFFFFFFu, active);
    if (!active) return INF_OK;
    BitReader br; br.init(words, byte_off, nbytes);
    HuffLim ll, dd;
#pragma unroll
    for (int i = 0; i < 15; i++) { ll.v[i] = 0; dd.v[i] = 0; }
    uint32_t pos = 0;                 // bytes produced so far (including match bytes still in flight)
    int state = ST_HDR, rc = INF_OK; uint32_t bfinal = 0, stored_rem = 0;
    // current match: bytes not yet loaded
    uint32_t m_rem = 0, m_dist = 0, m_wrap = 0xFFFFFFFFu, m_dst = 0; bool m_near = false;
    // loaded-but-not-yet-stored bytes of the previous round
    uint32_t pend_lo = 0, pend_hi = 0, pend_n = 0, pend_pos = 0;
    for (;;) {
        BD_STAT(g_inflate_stats.iters++);
        int sym = -1; bool have_sym = false;
        if (m_rem == 0) {
            if (state == ST_HDR) { rc = next_block(t, br, ll, dd, lens, pos, isize, state, bfinal, stored_rem); if (rc) state = ST_DONE; }
            if (state == ST_SYM) {
                br.refill(); sym = decode_sym<Tab, 0>(t, ll, br); have_sym = true;
                if (sym < 0) { rc = INF_ERR_CODE; state = ST_DONE; have_sym = false; }
            } else if (state == ST_STORED) {
                br.refill(); sym = (int)br.get(8); have_sym = true;
                if (--stored_rem == 0) state = bfinal ? ST_DONE : ST_HDR;
            }
        }
        // ---- A: store the bytes loaded in the previous round (their loads have had a whole decode to land)
        if (pend_n) {
            uint64_t a = a0 + pend_pos;
#pragma unroll
            for (int k = 0; k < 8; k++) if ((uint32_t)k < pend_n) t.ring_st8((uint32_t)(a + k), ((k < 4 ? pend_lo : pend_hi) >> ((k & 3) * 8)) & 0xFFu);
            uint64_t e = a + pend_n;
            if ((e >> 4) != (a >> 4)) flush_line(t, out, ((e >> 4) - 1) << 4, a0);
            pend_n = 0;
        }
        // ---- act on the decoded symbol
        if (have_sym) {
            if (sym < 256) {
                if (pos >= isize) { rc = INF_ERR_OVERRUN; state = ST_DONE; }
                else {
                    BD_STAT(g_inflate_stats.lits++);
                    uint64_t a = a0 + pos;
                    t.ring_st8((uint32_t)a, (uint32_t)sym);
                    if (((a + 1) & 15) == 0) flush_line(t, out, (a + 1) - 16, a0);
                    pos++;
                }
            } else if (sym == 256) {
                state = bfinal ? ST_DONE : ST_HDR;
            } else {
                uint32_t s = (uint32_t)sym - 257u;
                uint32_t len = 0, dist = 0; int ds = -1;
                if (s <= 28) {
                    if (s < 8) len = 3 + s;
                    else if (s == 28) len = 258;
                    else { uint32_t eb = (s >> 2) - 1; len = ((4 + (s & 3)) << eb) + 3 + br.get((int)eb); }
                    br.refill();
                    ds = decode_sym<Tab, 1>(t, dd, br);
                }
                if (ds >= 0 && ds <= 29) {
                    if (ds < 4) dist = 1 + (uint32_t)ds;
                    else { uint32_t eb = ((uint32_t)ds >> 1) - 1; dist = ((2 + ((uint32_t)ds & 1)) << eb) + 1 + br.get((int)eb); }
                }
                int e = INF_OK;
                if (s > 28 || ds < 0 || ds > 29) e = INF_ERR_CODE;
                else if (dist > pos) e = INF_ERR_DIST;
                else if (pos + len > isize) e = INF_ERR_OVERRUN;
                else if (br.overrun()) e = INF_ERR_INPUT;
                if (e) { rc = e; state = ST_DONE; }
                else {
                    BD_STAT(g_inflate_stats.matches++); BD_STAT(g_inflate_stats.match_bytes += len); BD_STAT(g_inflate_stats.mlen_hist[len >= 64 ? 15 : len / 4]++);
                    BD_STAT(g_inflate_stats.near_matches += dist <= NEAR_MAX);
                    m_dist = dist; m_wrap = dist < 8 ? dist : 0xFFFFFFFFu; m_near = dist <= NEAR_MAX;
                    m_dst = pos; m_rem = len; pos += len;
                }
            }
        }
        // ---- B: load one round (<= 8 bytes) of the current match; they are stored next iteration
        if (m_rem) {
            uint32_t n = m_rem < 8 ? m_rem : 8;
            // byte k of this round equals byte (k mod dist) of the dist bytes that precede the round, all of which
            // are already stored (the pattern is periodic with period dist); for dist >= 8 there is no wrap
            uint64_t sbase = a0 + m_dst - m_dist; uint32_t m_o = 0;
            pend_lo = 0; pend_hi = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if ((uint32_t)k < n) {
                    uint32_t b = m_near ? t.ring_ld8((uint32_t)(sbase + m_o)) : out.get(sbase + m_o);
                    if (k < 4) pend_lo |= b << (k * 8); else pend_hi |= b << ((k - 4) * 8);
                    m_o++; if (m_o == m_wrap) m_o = 0;
                }
            }
            pend_pos = m_dst; pend_n = n; m_dst += n; m_rem -= n;
        }
        bool cont = !(state == ST_DONE && m_rem == 0 && pend_n == 0);
        mask = BD_BALLOT(mask, cont);
        if (!cont) break;
    }
    if (rc) return rc;
    // flush the last, partial line (byte stores: the rest of the line belongs to the next block)
    {
        uint64_t e = a0 + pos, ls = e & ~15ull; if (ls < a0) ls = a0;
        for (uint64_t a = ls; a < e; a++) out.put(a, t.ring_ld8((uint32_t)a));
    }
    return pos == isize ? INF_OK : INF_ERR_SHORT;
}

}  // namespace bdk
