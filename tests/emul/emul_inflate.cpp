// Host emulation harness (TEST ONLY): runs the exact lane-level DEFLATE logic of
// sambamba_b200/csrc/inflate_core.cuh on the CPU, one "lane" at a time, so it can be checked
// against zlib without a GPU.  Never linked into libbdepth.so.
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define BD_INFLATE_STATS 1
#include "../../sambamba_b200/csrc/inflate_core.cuh"
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
