// launch.cuh -- the two spellings the rest of the sources use for things only nvcc understands.
//   BD_LAUNCH(grid, block, smem_bytes, stream, kernel<targs>)(args...)   ==  kernel<targs><<<grid, block, smem_bytes, stream>>>(args...)
//   BD_DYN_SMEM(T, name);                                                  ==  extern __shared__ T name[];
//   BD_NOINLINE                                                            ==  __noinline__
// Written as macros so that a test build (-DBDEPTH_EMULATE_SHIM='"<header>"', see tests/emul/Makefile) can compile the
// very same pipeline and kernels with g++ against a CUDA-on-CPU emulation and exercise launch plumbing without a GPU.
// The product is always built by nvcc without that define and contains nothing of the emulation.
#pragma once
#ifdef BDEPTH_EMULATE_SHIM
#include BDEPTH_EMULATE_SHIM
#else
#include <cuda_runtime.h>
#define BD_LAUNCH(g, b, s, st, ...) __VA_ARGS__<<<(g), (b), (s), (st)>>>
#define BD_DYN_SMEM(T, name) extern __shared__ T name[]
#define BD_NOINLINE __noinline__
#endif
