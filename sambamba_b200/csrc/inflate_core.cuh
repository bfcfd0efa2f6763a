// inflate_core.cuh -- raw DEFLATE (RFC 1951) decoder for one BGZF block, written so that ONE
// CUDA LANE decodes ONE block with its Huffman tables in (lane-interleaved) shared memory.
//
// Replaces: BioD/bio/core/bgzf/block.d:127-216 (decompressBgzfBlock -> zlib inflateInit2(-15)/
// inflate(Z_FINISH)/inflateEnd, bound through BioD/bio/core/utils/zlib.d:143-162).  The
// arithmetic being restated is zlib's (third-party, not under /root/reference); RFC 1951 output
// is deterministic so any correct inflater is bit-identical.
//
// Design (see DESIGN.md "K1"): a warp-per-block decoder issues ~100 instructions per symbol with
// one useful lane; giving every lane its own block makes those instructions decode 32 symbols.
// The price is per-lane table space, which bounds occupancy, so there is NO lookup table:
// decoding is canonical-Huffman arithmetic.  Measured on BAM data, a first-level LUT would miss
// on 5-10 % of symbols, i.e. in ~97 % of 32-lane iterations, so its fast path would almost never
// save the warp anything while costing 4x the shared memory.
//   per code: left-justified 15-bit limits lim[1..15] live in REGISTERS
//             len  = 1 + #{ j in 1..14 : peek15 >= lim[j] }         (branch-free compares)
//             sym  = sorted[ (peek15 >> (15-len)) + delta[len] ]     (two shared-memory loads)
//   litlen : 288 x u16 sorted symbols (144 words) + 15 x i16 delta (8 words)
//   dist   :  32 x u8  sorted symbols (  8 words) + 15 x i16 delta (8 words)
// = 168 words (672 B) per lane, 21 KB per warp, 10 warps per SM.
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
struct InflateStats { unsigned long long lits, matches, match_bytes, tables, len_hist[16], mlen_hist[16]; };
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
constexpr int T_WORDS = 168;

// ---- table storage policies -------------------------------------------------------------
// Lane-interleaved shared memory: word w of this lane lives at base[w * 32]; every lane always
// hits its own bank, so data-dependent indices never conflict.
struct SmemTab {
    uint32_t* base;
    BD_HD uint32_t ldw(int w) const { return base[w * 32]; }
    BD_HD void stw(int w, uint32_t v) const { base[w * 32] = v; }
};
// Plain array (host tests).
struct FlatTab {
    uint32_t* base;
    BD_HD uint32_t ldw(int w) const { return base[w]; }
    BD_HD void stw(int w, uint32_t v) const { base[w] = v; }
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

// ---- bit reader: 64-bit reservoir fed by aligned 32-bit words -----------------------------
struct BitReader {
    const uint32_t* wp;
    const uint32_t* wend;   // one past the last word that may contain stream bits
    uint64_t bb;
    int bc;
    BD_HD void init(const uint32_t* words, uint64_t byte_off, uint32_t nbytes) {
        wp = words + (byte_off >> 2);
        wend = words + ((byte_off + nbytes + 3) >> 2);
        unsigned mis = (unsigned)(byte_off & 3);
        bb = (uint64_t)(*wp++) >> (8 * mis);
        bc = 32 - 8 * (int)mis;
    }
    BD_HD void refill() {   // afterwards bc >= 33
        if (bc <= 32) { bb |= (uint64_t)(*wp++) << bc; bc += 32; }
    }
    BD_HD uint32_t peek(int n) const { return (uint32_t)bb & ((1u << n) - 1u); }
    BD_HD void drop(int n) { bb >>= n; bc -= n; }
    BD_HD uint32_t get(int n) { uint32_t v = peek(n); drop(n); return v; }
    // the reservoir prefetches at most two words past the last stream word
    BD_HD bool overrun() const { return wp > wend + 2; }
};

// Left-justified 15-bit limits, index len-1.  lim[14] == 32768 for a complete code.
struct HuffLim { uint32_t v[15]; };

// Build one canonical table from code lengths lens[0..n).  KIND 0 = litlen (u16 symbols),
// 1 = dist (u8 symbols).  Mirrors zlib inflate_table()'s validity rules: over-subscribed ->
// error; incomplete -> error unless the longest code is <= 1 bit (dist may also be empty).
template <class Tab, int KIND>
BD_HD int build_table(const Tab& t, const uint8_t* lens, int n, HuffLim& lim) {
    uint32_t cnt[16];
    for (int i = 0; i < 16; i++) cnt[i] = 0;
    for (int s = 0; s < n; s++) cnt[lens[s]]++;
    int left = 1, maxl = 0;
    for (int l = 1; l <= 15; l++) { left <<= 1; left -= (int)cnt[l]; if (left < 0) return INF_ERR_TABLE; if (cnt[l]) maxl = l; }
    if (left > 0 && (KIND == 0 ? maxl != 1 : maxl > 1)) return INF_ERR_TABLE;
    uint32_t nxt[16];      // next free slot in the sorted table per length
    uint32_t code = 0, o = 0;
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
    int L = 1;
#pragma unroll
    for (int j = 0; j < 14; j++) L += (rev15 >= lim.v[j]) ? 1 : 0;
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

// Output policy: plain byte pointer (global memory on the device).
struct ByteOut {
    uint8_t* p;
    BD_HD void put(uint32_t i, uint32_t v) const { p[i] = (uint8_t)v; }
    BD_HD uint32_t get(uint32_t i) const { return p[i]; }
};

// Inflate one raw-deflate stream.  words/byte_off/nbytes locate the compressed data inside a
// 4-byte-aligned buffer that has >= 64 readable bytes after the last block.
template <class Tab, class Out>
BD_HD int inflate_block(const Tab& t, const uint32_t* words, uint64_t byte_off, uint32_t nbytes, const Out& out, uint32_t isize, uint8_t* lens /* 320 B scratch */) {
    BitReader br; br.init(words, byte_off, nbytes);
    HuffLim ll, dd;
    uint32_t pos = 0;
    for (;;) {
        br.refill();
        uint32_t bfinal = br.get(1), btype = br.get(2);
        if (btype == 0) {
            br.drop(br.bc & 7);
            br.refill();
            uint32_t len = br.get(16); br.refill(); uint32_t nlen = br.get(16);
            if ((len ^ 0xFFFFu) != nlen) return INF_ERR_STORED;
            if (pos + len > isize) return INF_ERR_OVERRUN;
            for (uint32_t i = 0; i < len; i++) { br.refill(); out.put(pos++, br.get(8)); }
            if (br.overrun()) return INF_ERR_INPUT;
        } else if (btype == 3) {
            return INF_ERR_BTYPE;
        } else {
            int nl, nd;
            if (btype == 1) {
                // fixed code (RFC 1951 3.2.6): complete over 288 litlen / 32 distance symbols; litlen 286/287 and
                // distance 30/31 are rejected after decoding, as zlib does
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
            for (;;) {
                br.refill();
                int sym = decode_sym<Tab, 0>(t, ll, br);
                if (sym < 256) {
                    if (sym < 0) return INF_ERR_CODE;
                    if (pos >= isize) return INF_ERR_OVERRUN;
                    BD_STAT(g_inflate_stats.lits++);
                    out.put(pos++, (uint32_t)sym);
                    continue;
                }
                if (sym == 256) break;
                uint32_t s = (uint32_t)sym - 257u;
                if (s > 28) return INF_ERR_CODE;
                uint32_t len;
                if (s < 8) len = 3 + s;
                else if (s == 28) len = 258;
                else { uint32_t eb = (s >> 2) - 1; len = ((4 + (s & 3)) << eb) + 3 + br.get((int)eb); }
                br.refill();
                int ds = decode_sym<Tab, 1>(t, dd, br);
                if (ds < 0 || ds > 29) return INF_ERR_CODE;
                uint32_t dist;
                if (ds < 4) dist = 1 + (uint32_t)ds;
                else { uint32_t eb = ((uint32_t)ds >> 1) - 1; dist = ((2 + ((uint32_t)ds & 1)) << eb) + 1 + br.get((int)eb); }
                if (dist > pos) return INF_ERR_DIST;
                if (pos + len > isize) return INF_ERR_OVERRUN;
                BD_STAT(g_inflate_stats.matches++); BD_STAT(g_inflate_stats.match_bytes += len); BD_STAT(g_inflate_stats.mlen_hist[len >= 64 ? 15 : len / 4]++);
                // LZ77 copy in rounds of 4: all loads of a round are independent of its stores.
                // For dist < 4 the source index wraps inside the dist-byte pattern.
                uint32_t src = pos - dist, o = 0, wrap = dist < 4 ? dist : 0xFFFFFFFFu;
                for (uint32_t i = 0; i < len; i += 4) {
                    uint32_t b[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) { b[k] = (i + k < len) ? out.get(src + o) : 0; o++; if (o == wrap) o = 0; }
#pragma unroll
                    for (int k = 0; k < 4; k++) if (i + k < len) out.put(pos + i + k, b[k]);
                }
                pos += len;
                if (br.overrun()) return INF_ERR_INPUT;
            }
        }
        if (bfinal) break;
        if (br.overrun()) return INF_ERR_INPUT;
    }
    return pos == isize ? INF_OK : INF_ERR_SHORT;
}

}  // namespace bdk
