// Host emulation harness (TEST ONLY): runs the exact lane-level DEFLATE logic of
// sambamba_b200/csrc/inflate_core.cuh on the CPU, one "lane" at a time, so it can be checked
// against zlib without a GPU.  Never linked into libbdepth.so.
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define BD_INFLATE_STATS 1
#include "../../sambamba_b200/csrc/inflate_core.cuh"
#include "../../sambamba_b200/csrc/inflate2_core.cuh"
namespace bdk { InflateStats g_inflate_stats; }
using namespace bdk;

extern "C" void emul_stats(unsigned long long* o) { memcpy(o, &g_inflate_stats, sizeof(g_inflate_stats)); memset(&g_inflate_stats, 0, sizeof(g_inflate_stats)); }

static bool g_lit3 = false;
extern "C" void emul_set_lit3(int on) { g_lit3 = on != 0; }       // the three-literal variant of the lane logic (k1_inflate_lit3)
extern "C" long emul_inflate_file(const char* path, uint8_t* dst, uint64_t cap, int* first_err) {
    FILE* f = fopen(path, "rb"); if (!f) return -1;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint32_t> words((n + 64 + 3) / 4 + 16, 0);
    if (fread(words.data(), 1, n, f) != (size_t)n) { fclose(f); return -1; }
    fclose(f);
    const uint8_t* file = (const uint8_t*)words.data();
    uint64_t off = 0, uoff = 0; *first_err = 0;
    uint32_t tab[T_WORDS]; uint32_t lens[96];
    while (off + 18 <= (uint64_t)n) {
        const uint8_t* p = file + off;
        uint32_t xlen = p[10] | (p[11] << 8), bsize = 0;
        for (uint32_t l = 0; l < xlen;) { uint32_t slen = p[14 + l] | (p[15 + l] << 8); if (p[12 + l] == 66 && p[13 + l] == 67) bsize = p[16 + l] | (p[17 + l] << 8); l += 4 + slen; }
        uint32_t total = bsize + 1, cdata = total - xlen - 20;
        uint32_t isize = p[total - 4] | (p[total - 3] << 8) | (p[total - 2] << 16) | ((uint32_t)p[total - 1] << 24);
        if (isize == 0) break;
        if (uoff + isize > cap) return -2;
        memset(tab, 0xAB, sizeof tab);
        FlatTab ft{tab}; ByteOut out{dst};
        int rc = g_lit3 ? inflate_block<FlatTab, ByteOut, true>(ft, words.data(), off + 12 + xlen, cdata, out, uoff, isize, lens) : inflate_block(ft, words.data(), off + 12 + xlen, cdata, out, uoff, isize, lens);
        if (rc) { *first_err = rc; return -(100 + rc); }
        uoff += isize; off += total;
    }
    return (long)uoff;
}
// decode a single raw deflate stream (for crafted-stream tests); byte_off exercises misaligned starts
extern "C" int emul_inflate_raw(const uint8_t* src, uint32_t n, uint8_t* dst, uint32_t isize, uint32_t byte_off, uint32_t out_off) {
    std::vector<uint32_t> words((n + byte_off + 64 + 3) / 4 + 16, 0);
    memcpy((uint8_t*)words.data() + byte_off, src, n);
    uint32_t tab[T_WORDS]; uint32_t lens[96]; memset(tab, 0xCD, sizeof tab);
    FlatTab ft{tab}; ByteOut out{dst};
    return g_lit3 ? inflate_block<FlatTab, ByteOut, true>(ft, words.data(), byte_off, n, out, out_off, isize, lens) : inflate_block(ft, words.data(), byte_off, n, out, out_off, isize, lens);
}

// ---- the two-phase decoder (inflate2_core.cuh): phase 1 as one lane, phase 2 in its serial restatement
static bool g_lims = false;
extern "C" void emul_set_lims(int on) { g_lims = on != 0; }        // Huffman limits in table storage (LIMS) or in registers
static int two_phase_block(const uint32_t* words, uint64_t byte_off, uint32_t nbytes, uint8_t* dst, uint64_t a0, uint32_t isize, uint32_t tok_cap, unsigned long long* stats) {
    uint32_t tab[H_WORDS + 8]; uint32_t lens[96]; memset(tab, 0xAB, sizeof tab);
    std::vector<uint32_t> tok(tok_cap + 1, 0xDDDDDDDDu); std::vector<uint8_t> lit(MAX_SEG * 256, 0xEE), lits(((size_t)isize + 15) / 16 * 16 + 16, 0xCC);
    uint32_t seg[MAX_SEG], ntok = 0, nseg = 0, nlit = 0;
    uint32_t lims[16]; memset(lims, 0xCD, sizeof lims);
    FlatTab2 ft{tab, lims};
    HuffOut ho{tok.data(), tok_cap, lits.data(), lit.data(), seg};
    int rc = g_lims ? huff_phase<FlatTab2, true, 4>(ft, words, byte_off, nbytes, isize, lens, ho, ntok, nseg, nlit)
                    : huff_phase<FlatTab2, false, 4>(ft, words, byte_off, nbytes, isize, lens, ho, ntok, nseg, nlit);
    if (stats) { stats[0] += ntok; stats[1] += nseg; if (rc == INF_FALLBACK) stats[2]++; stats[3] += nlit; }
    if (rc) return rc;
    if (tok[tok_cap] != 0xDDDDDDDDu) return -77;      // wrote past the token area
    return lz_phase_serial(dst + a0, isize, tok.data(), ntok, lits.data(), nlit, lit.data(), seg, nseg);
}
extern "C" long emul_inflate2_file(const char* path, uint8_t* dst, uint64_t cap, int* first_err, unsigned long long* stats /* tokens, segments, fallbacks, literals */, uint32_t tok_cap_div) {
    FILE* f = fopen(path, "rb"); if (!f) return -1;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint32_t> words((n + 64 + 3) / 4 + 16, 0);
    if (fread(words.data(), 1, n, f) != (size_t)n) { fclose(f); return -1; }
    fclose(f);
    const uint8_t* file = (const uint8_t*)words.data();
    uint64_t off = 0, uoff = 0; *first_err = 0;
    while (off + 18 <= (uint64_t)n) {
        const uint8_t* p = file + off;
        uint32_t xlen = p[10] | (p[11] << 8), bsize = 0;
        for (uint32_t l = 0; l < xlen;) { uint32_t slen = p[14 + l] | (p[15 + l] << 8); if (p[12 + l] == 66 && p[13 + l] == 67) bsize = p[16 + l] | (p[17 + l] << 8); l += 4 + slen; }
        uint32_t total = bsize + 1, cdata = total - xlen - 20;
        uint32_t isize = p[total - 4] | (p[total - 3] << 8) | (p[total - 2] << 16) | ((uint32_t)p[total - 1] << 24);
        if (isize == 0) break;
        if (uoff + isize > cap) return -2;
        int rc = two_phase_block(words.data(), off + 12 + xlen, cdata, dst, uoff, isize, isize / tok_cap_div + 16, stats);
        if (rc) { *first_err = rc; return -(100 + rc); }
        uoff += isize; off += total;
    }
    return (long)uoff;
}
extern "C" int emul_inflate2_raw(const uint8_t* src, uint32_t n, uint8_t* dst, uint32_t isize, uint32_t byte_off, uint32_t out_off) {
    std::vector<uint32_t> words((n + byte_off + 64 + 3) / 4 + 16, 0);
    memcpy((uint8_t*)words.data() + byte_off, src, n);
    return two_phase_block(words.data(), byte_off, n, dst, out_off, isize, isize / 3 + 300, nullptr);
}
