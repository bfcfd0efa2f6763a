// Self-test of the CUDA-on-CPU emulation (TEST INFRASTRUCTURE): small kernels with known answers for every collective the
// library's kernels use -- the emulation carries the CPU verification of the pipeline, so it is checked on its own.
#include "cuda_shim.hpp"
#include <stdio.h>
#include <vector>

__global__ void k_collectives(uint32_t* out, unsigned long long* sum, uint32_t* ors) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    uint64_t s = 1000 + lane;
    uint64_t s0 = __shfl_sync(0xFFFFFFFFu, s, 0);                                   // 1000
    unsigned odd = __ballot_sync(0xFFFFFFFFu, lane & 1);                            // 0xAAAAAAAA
    uint32_t up = __shfl_up_sync(0xFFFFFFFFu, lane, 1);                             // lane - 1 (lane 0 keeps its own)
    unsigned long long red = lane;
    for (int d = 16; d; d >>= 1) red += __shfl_xor_sync(0xFFFFFFFFu, red, d);       // 496 in every lane
    uint32_t orv = __reduce_or_sync(0xFFFFFFFFu, 1u << (lane & 7));                 // 0xFF
    uint32_t mn = __reduce_min_sync(0xFFFFFFFFu, 100u - lane);                      // 69
    int all = __all_sync(0xFFFFFFFFu, lane < 32), none = __all_sync(0xFFFFFFFFu, lane < 31);
    // a loop whose trip count differs per lane group but whose collectives are executed by everybody
    unsigned m = __ballot_sync(0xFFFFFFFFu, (lane % 5) == 0); uint32_t acc = 0;
    while (m) { int b = __ffs(m) - 1; m &= m - 1; acc += __shfl_sync(0xFFFFFFFFu, lane * 3u, b); }    // 3 * (0+5+...+30) = 315
    out[tid] = (uint32_t)(s0 == 1000) + 2u * (odd == 0xAAAAAAAAu) + 4u * (up == (lane ? lane - 1 : 0)) + 8u * (red == 496) + 16u * (orv == 0xFF) + 32u * (mn == 69)
             + 64u * (all == 1 && none == 0) + 128u * (acc == 315);
    atomicAdd(sum, (unsigned long long)tid);
    atomicOr(ors, 1u << (tid & 31));
}
__global__ void k_block(uint32_t* out) {          // __syncthreads + static shared memory + early exit of part of the block
    __shared__ uint32_t part[8];
    BD_DYN_SMEM(uint32_t, dyn);
    uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x >= 224) return;                 // the last warp leaves before the barrier
    uint32_t v = threadIdx.x;
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    if (lane == 0) { part[warp] = v; dyn[warp] = v * 2; }
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0, u = 0; for (int w = 0; w < 7; w++) { t += part[w]; u += dyn[w]; } out[blockIdx.x] = t + (u == 2 * t ? 0 : 1000000); }
}
__global__ void k_masked(uint32_t* out, uint32_t n) {     // the K1 pattern: a shrinking mask of lanes still running
    uint32_t lane = threadIdx.x & 31; unsigned mask = __ballot_sync(0xFFFFFFFFu, lane < n); if (lane >= n) return;
    uint32_t rounds = 0;
    for (;;) { rounds++; bool cont = rounds <= lane; mask = __ballot_sync(mask, cont); if (!cont) break; }
    out[lane] = rounds * 100 + __popc(mask);       // lane l leaves in round l+1 and sees the lanes above it still running
}

int main() {
    int bad = 0;
    { const unsigned G = 5, Bk = 256; std::vector<uint32_t> out(G * Bk, 0); unsigned long long sum = 0; uint32_t ors = 0;
      BD_LAUNCH(G, Bk, 0, nullptr, k_collectives)(out.data(), &sum, &ors);
      for (unsigned i = 0; i < G * Bk; i++) if (out[i] != 255) { if (bad < 5) printf("collectives: thread %u got %u\n", i, out[i]); bad++; }
      if (sum != (unsigned long long)(G * Bk) * (G * Bk - 1) / 2 || ors != 0xFFFFFFFFu) { printf("atomics: %llu %x\n", sum, ors); bad++; } }
    { std::vector<uint32_t> out(3, 0); BD_LAUNCH(3, 256, 64, nullptr, k_block)(out.data());
      for (int b = 0; b < 3; b++) if (out[b] != 223u * 224u / 2u) { printf("block %d: %u\n", b, out[b]); bad++; } }
    { std::vector<uint32_t> out(32, 0); BD_LAUNCH(1, 32, 0, nullptr, k_masked)(out.data(), 20u);
      for (uint32_t l = 0; l < 20; l++) if (out[l] != (l + 1) * 100 + (19 - l)) { printf("masked lane %u: %u\n", l, out[l]); bad++; } }
    printf(bad ? "FAILED %d\n" : "ok\n", bad);
    return bad ? 1 : 0;
}
