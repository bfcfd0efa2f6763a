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
                          uint32_t* seg_reads, uint32_t* seg_mbases, uint32_t n_samples_out, int force_general, int order, int* err2, unsigned long long* stat3,
                          const uint64_t* seg_u, const uint64_t* seg_qmin, uint64_t seg_ext_max) {
    std::vector<uint64_t> mhash(R ? R : 1); std::vector<uint32_t> mflag(R ? R : 1);
    err2[0] = err2[1] = 0; stat3[0] = stat3[1] = stat3[2] = 0;
    // K3's read index as k3_tile_index / k2_decode build it (kernels.cuh): passing short reads per 1024-position tile, long-read list
    uint64_t gmin = ~0ull, gmax = 0;
    for (uint32_t r = 0; r < R; r++) if (meta[r] & 1u) { gmin = std::min(gmin, start[r]); gmax = std::max(gmax, start[r] + span[r]); }
    uint64_t tiles_base = gmin == ~0ull ? 0 : gmin / 1024 * 1024; uint32_t n_tiles = gmin == ~0ull ? 0 : (uint32_t)((gmax - tiles_base + 1023) / 1024);
    std::vector<uint32_t> tile_lo(n_tiles + 2, 0xFFFFFFFFu), long_list;
    for (uint32_t r = 0; r < R; r++) {
        if ((meta[r] & 3u) == 3u) long_list.push_back(r);
        if ((meta[r] & 3u) == 1u) for (uint64_t t = (start[r] - tiles_base) / 1024; t <= (start[r] + span[r] - 1 - tiles_base) / 1024 && t < n_tiles; t++) tile_lo[t] = std::min(tile_lo[t], r);
    }
    MateParams p{start, span, meta, off, ncl, lseq, u, R, mhash.data(), mflag.data(), flt_s, flt_e, n_flt, counts, cnt_base, win_len, S, minq,
                 seg_s, seg_e, seg_pmax, seg_id, n_seg, seg_reads, seg_mbases, n_samples_out, seg_u, seg_qmin, seg_ext_max,
                 tile_lo.data(), tiles_base, n_tiles, long_list.data(), (uint32_t)long_list.size(),
                 0u, ~0ull, 0ull, 0ull, 1, nullptr, nullptr, 0, nullptr, 0u, INT64_MIN, force_general, err2, stat3};      // one batch, one rank: no ghosts, everything closes
    std::vector<uint32_t> ord(R); std::iota(ord.begin(), ord.end(), 0u);
    if (order == 1) std::reverse(ord.begin(), ord.end());
    if (order == 2) { uint64_t s = 88172645463325252ull; for (uint32_t i = R; i > 1; i--) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; std::swap(ord[i - 1], ord[s % i]); } }
    for (uint32_t r : ord) mate_hash_one(p, r);
    for (uint32_t r : ord) mate_link_one(p, r);
    for (uint32_t r : ord) mate_fix_one(p, r);
    return err2[0];
}
