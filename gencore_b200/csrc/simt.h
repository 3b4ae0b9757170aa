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
#define GCB_GRID_DEP() ((void)0)
#else
#include <cuda_runtime.h>
#include <string.h>
// -DGCB_PDL=1: every kernel of the path is launched with programmatic stream serialization (programmatic dependent launch):
// a kernel lets its successor be scheduled at once (griddepcontrol.launch_dependents, first thing) and itself waits for its
// predecessor's completion and memory flush before it touches anything (griddepcontrol.wait: GCB_GRID_DEP() at the top of every
// kernel).  Measured on the B200 it LOSES 3 % on the BASELINE shape (0.492 against 0.476 ms per pass: the early-resident CTAs
// of the next kernel take warp slots from the tail of the running one, and the dozen launches of a batch were only 3 % of the
// pass to begin with), so plain launches are the default (profiles/r03_notes.md).
#if defined(GCB_PDL) && GCB_PDL
template <typename... KArgs, typename... Args>
static inline void gcb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at;
    memset(&at, 0, sizeof at);
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define GCB_LAUNCH(kernel, grid, block, smem, stream, ...) gcb_launch(kernel, (grid), (block), (smem), (cudaStream_t)(stream), __VA_ARGS__)
#define GCB_GRID_DEP() asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory")
#else
#define GCB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define GCB_GRID_DEP() ((void)0)
#endif
#define GCB_DYN_SMEM(name) extern __shared__ __align__(128) uint8_t name[]
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "gencore_b200 is written for sm_100a (B200) only"
#endif
#endif

#define GCB_HD __host__ __device__ __forceinline__
#define GCB_DEV __device__ __forceinline__
