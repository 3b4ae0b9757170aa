// simt.h — the one place where the build mode is decided.
//
//   product build (nvcc, sm_100a):   the CUDA runtime and the real intrinsics.
//   GCB_SIMT_CHECK (g++, tests only): tests/simt_check/simt_check.h supplies a single-threaded SIMT
//       interpreter (fibers, warp collectives, block barriers) so the SAME kernel source can be
//       executed on the CPU-only CI box and compared bit-for-bit with the oracle.  That build is
//       made by tests/simt_check/build.py into tests/simt_check/_build/ and is loaded by tests only;
//       nothing under gencore_b200/ ever loads it — the product has no CPU path.
#pragma once

#ifdef GCB_SIMT_CHECK
#include "simt_check.h"
#define GCB_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                    \
        if (getenv("GCB_SIMT_TRACE")) fprintf(stderr, "simt launch %s\n", #kernel);         \
        ::simt::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); });            \
    } while (0)
#define GCB_DYN_SMEM(name) uint8_t *name = ::simt::dyn_smem()
#else
#include <cuda_runtime.h>
#define GCB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define GCB_DYN_SMEM(name) extern __shared__ __align__(128) uint8_t name[]
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "gencore_b200 is written for sm_100a (B200) only"
#endif
#endif

#define GCB_HD __host__ __device__ __forceinline__
#define GCB_DEV __device__ __forceinline__
