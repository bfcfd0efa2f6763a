// Host emulation harness (TEST ONLY): runs the thread bodies of sambamba_b200/csrc/mates.cuh (km_hash, km_link,
// km_fix) on the CPU, one emulated thread after another in a chosen order, over a record SoA laid out exactly as
// k2_decode writes it.  Never linked into libbdepth.so.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>
#include "../../sambamba_b200/csrc/mates.cuh"
using namespace bdk;

extern "C" int emul_mates(const uint8_t* u, uint32_t R, const uint64_t* start, const uint32_t* span, const uint32_t* meta, const int64_t* off,
                          const uint32_t* ncl, const int32_t* lseq, const uint64_t* flt_s, const uint64_t* flt_e, uint32_t n_flt,
                          uint32_t* counts, uint64_t cnt_base, uint64_t win_len, uint32_t S, uint32_t minq,
                          const uint64_t* seg_s, const uint64_t* seg_e, const uint64_t* seg_pmax, const uint32_t* seg_id, uint32_t n_seg,
                          uint32_t* seg_reads, uint32_t* seg_mbases, uint32_t n_samples_out, int force_general, int order, int* err2, unsigned long long* stat3) {
    std::vector<uint64_t> mhash(R ? R : 1); std::vector<uint32_t> mflag(R ? R : 1);
    err2[0] = err2[1] = 0; stat3[0] = stat3[1] = stat3[2] = 0;
    MateParams p{start, span, meta, off, ncl, lseq, u, R, mhash.data(), mflag.data(), flt_s, flt_e, n_flt, counts, cnt_base, win_len, S, minq,
                 seg_s, seg_e, seg_pmax, seg_id, n_seg, seg_reads, seg_mbases, n_samples_out, force_general, err2, stat3};
    std::vector<uint32_t> ord(R); std::iota(ord.begin(), ord.end(), 0u);
    if (order == 1) std::reverse(ord.begin(), ord.end());
    if (order == 2) { uint64_t s = 88172645463325252ull; for (uint32_t i = R; i > 1; i--) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; std::swap(ord[i - 1], ord[s % i]); } }
    for (uint32_t r : ord) mate_hash_one(p, r);
    for (uint32_t r : ord) mate_link_one(p, r);
    for (uint32_t r : ord) mate_fix_one(p, r);
    return err2[0];
}
