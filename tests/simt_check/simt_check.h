// simt_check.h — a small single-threaded SIMT interpreter for CPU-only CI.  TEST INFRASTRUCTURE ONLY.
//
// The kernels under gencore_b200/csrc/ are compiled a second time by g++ with -DGCB_SIMT_CHECK
// (tests/simt_check/build.py) against this header: every CUDA thread of a block becomes a ucontext
// fiber, __syncthreads()/warp collectives are rendezvous points between fibers, blocks run one
// after another, and the CUDA runtime calls the host side makes become malloc/memcpy.  That lets the
// parity tests drive the real kernel source against the oracle on a box without a GPU; it proves
// nothing about races, alignment or speed (the `-m gpu` tests and compute-sanitizer do that) and it
// is never loaded by the product package.
#pragma once

#include <ucontext.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(8))) int2 { int x, y; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 v = {x, y}; return v; }
inline int2 make_int2(int x, int y) { int2 v = {x, y}; return v; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace simt {

constexpr int kWarp = 32;
constexpr size_t kStackBytes = 256 * 1024;

struct Fiber {
    ucontext_t ctx;
    char *stack = nullptr;
    bool done = true;
};

struct SubGroup {        // the lanes of one partial mask (keyed by the mask itself: masks that share lanes may be in use at once)
    unsigned mask = 0;
    uint64_t xchg[2][kWarp];
    int parity = 0;
    int arrived = 0;
    unsigned generation = 0;
};

struct WarpState {
    uint64_t xchg[2][kWarp];
    int parity = 0;       // which xchg buffer the next collective uses
    int arrived = 0;
    unsigned generation = 0;
    int live = 0;
    std::vector<SubGroup> sub;
    SubGroup &group(unsigned mask) {
        for (SubGroup &g : sub)
            if (g.mask == mask) return g;
        sub.reserve(64);  // (references handed out earlier must stay valid)
        if (sub.size() >= 64) {
            fprintf(stderr, "simt_check: more than 64 distinct partial masks in one warp\n");
            abort();
        }
        sub.emplace_back();
        sub.back().mask = mask;
        return sub.back();
    }
};

struct State {
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    ucontext_t sched;
    int cur = -1;
    int nthreads = 0;
    int live = 0;
    int bar_arrived = 0;
    unsigned bar_generation = 0;
    uint64_t progress = 0;
    uint64_t soft_progress = 0;  // polls of a spin wait (simt::relax): a fiber moved, but nothing was completed
    std::vector<uint8_t> smem;
    const std::function<void()> *body = nullptr;
    uint64_t launches = 0;
    int tags[2048] = {0};  // debugging aid: where every fiber last said it was (GCB_TRACE)
};

inline State &st() {
    static State s;
    return s;
}

}  // namespace simt

inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

namespace simt {

inline uint8_t *dyn_smem() { return st().smem.data(); }
inline void trace(int tag) { st().tags[st().cur] = tag; }
inline void relax();

inline void yield() {
    State &s = st();
    swapcontext(&s.fibers[s.cur].ctx, &s.sched);
}

inline void relax() {  // one poll of a spin wait
    st().soft_progress++;
    yield();
}

inline void trampoline() {
    State &s = st();
    (*s.body)();
    int me = s.cur;
    s.fibers[me].done = true;
    s.live--;
    s.warps[me / kWarp].live--;
    s.progress++;
    swapcontext(&s.fibers[me].ctx, &s.sched);
}

inline void run_block(int nthreads, const std::function<void()> &body) {
    State &s = st();
    if ((int)s.fibers.size() < nthreads) s.fibers.resize(nthreads);
    s.warps.assign((nthreads + kWarp - 1) / kWarp, WarpState());
    s.nthreads = nthreads;
    s.live = nthreads;
    s.bar_arrived = 0;
    s.body = &body;
    for (int t = 0; t < nthreads; t++) {
        Fiber &f = s.fibers[t];
        if (!f.stack) f.stack = (char *)malloc(kStackBytes);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStackBytes;
        f.ctx.uc_link = &s.sched;
        makecontext(&f.ctx, (void (*)())trampoline, 0);
        f.done = false;
        s.warps[t / kWarp].live++;
    }
    int idle_sweeps = 0;
    while (s.live > 0) {
        uint64_t before = s.progress, soft_before = s.soft_progress;
        for (int t = 0; t < nthreads; t++) {
            if (s.fibers[t].done) continue;
            s.cur = t;
            threadIdx.x = (unsigned)t;
            threadIdx.y = threadIdx.z = 0;
            swapcontext(&s.sched, &s.fibers[t].ctx);
        }
        // a sweep in which nothing completed is a deadlock unless fibers are polling (a poll loop interleaved with warp
        // collectives can go a few sweeps without completing one); polling alone for thousands of sweeps is one too
        idle_sweeps = s.progress == before ? idle_sweeps + 1 : 0;
        if (s.live > 0 && s.progress == before && (s.soft_progress == soft_before || idle_sweeps > 100000)) {
            fprintf(stderr, "simt_check: deadlock in block %u (%d threads alive, none can advance)\n", blockIdx.x, s.live);
            for (int q = 0; q < 64; q++) fprintf(stderr, "%s%016llx", q % 4 ? " " : "\n  smem ", (unsigned long long)((uint64_t *)s.smem.data())[q]);
            fprintf(stderr, "\n");
            for (int t = 0; t < nthreads; t += 1)
                if (!s.fibers[t].done && (t % kWarp == 0 || s.tags[t] != s.tags[t - 1])) fprintf(stderr, "  thread %d tag %d\n", t, s.tags[t]);
            abort();
        }
    }
}

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body) {
    State &s = st();
    if (block.x % kWarp != 0 || block.y != 1 || block.z != 1 || grid.y != 1 || grid.z != 1) {
        fprintf(stderr, "simt_check: unsupported launch shape\n");
        abort();
    }
    s.launches++;
    if (s.smem.size() < smem_bytes + 128) s.smem.resize(smem_bytes + 128);
    blockDim = block;
    gridDim = grid;
    for (unsigned b = 0; b < grid.x; b++) {
        blockIdx.x = b;
        blockIdx.y = blockIdx.z = 0;
        run_block((int)block.x, body);
    }
}

// ---- rendezvous primitives
inline void block_barrier() {
    State &s = st();
    unsigned gen = s.bar_generation;
    s.bar_arrived++;
    for (;;) {
        if (s.bar_generation != gen) return;
        if (s.bar_arrived >= s.live) {  // exited threads count as arrived (CUDA semantics)
            s.bar_arrived = 0;
            s.bar_generation++;
            s.progress++;
            return;
        }
        yield();
    }
}

inline void warp_barrier() {
    State &s = st();
    WarpState &w = s.warps[s.cur / kWarp];
    unsigned gen = w.generation;
    w.arrived++;
    for (;;) {
        if (w.generation != gen) return;
        if (w.arrived >= w.live) {
            if (w.live != kWarp) {
                fprintf(stderr, "simt_check: warp collective after %d lanes of the warp exited\n", kWarp - w.live);
                abort();
            }
            w.arrived = 0;
            w.generation++;
            s.progress++;
            return;
        }
        yield();
    }
}

inline int lane() { return st().cur % kWarp; }

// a collective over the lanes of a partial mask: all of them must be alive and must name the same mask
inline SubGroup &sub_barrier(unsigned mask) {
    State &s = st();
    const int wbase = (s.cur / kWarp) * kWarp;
    WarpState &w = s.warps[s.cur / kWarp];
    if (!((mask >> (s.cur % kWarp)) & 1u)) {
        fprintf(stderr, "simt_check: lane %d calls a collective whose mask %08x does not name it\n", s.cur % kWarp, mask);
        abort();
    }
    int need = 0;
    for (int l = 0; l < kWarp; l++)
        if ((mask >> l) & 1u) {
            if (wbase + l >= s.nthreads || s.fibers[wbase + l].done) {
                fprintf(stderr, "simt_check: collective mask %08x names lane %d, which has exited\n", mask, l);
                abort();
            }
            need++;
        }
    SubGroup &g = w.group(mask);
    unsigned gen = g.generation;
    g.arrived++;
    for (;;) {
        if (g.generation != gen) return g;
        if (g.arrived >= need) {
            g.arrived = 0;
            g.generation++;
            s.progress++;
            return g;
        }
        yield();
    }
}

// every lane publishes v, gets the whole vector back
inline const uint64_t *exchange(uint64_t v, unsigned mask = 0xffffffffu) {
    State &s = st();
    WarpState &w = s.warps[s.cur / kWarp];
    if (mask != 0xffffffffu) {
        SubGroup &g0 = w.group(mask);
        int p = g0.parity;
        g0.xchg[p][s.cur % kWarp] = v;
        SubGroup &g = sub_barrier(mask);
        if (g.parity == p) g.parity = p ^ 1;
        return g.xchg[p];
    }
    int p = w.parity;
    w.xchg[p][s.cur % kWarp] = v;
    warp_barrier();
    // the last arriver flips parity for the next collective; all lanes of this one still read buffer p
    if (w.parity == p) w.parity = p ^ 1;  // flipped once per collective, by the first lane released
    return w.xchg[p];
}

inline void check_full(unsigned) {}  // partial masks rendezvous among their own lanes (sub_barrier)

template <typename T>
inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "collective operand too wide");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T>
inline T from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

}  // namespace simt

// ---- CUDA device intrinsics used by the kernels ---------------------------------------------------
inline void __syncthreads() { simt::block_barrier(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) {
    if (mask != 0xffffffffu) simt::sub_barrier(mask);
    else simt::warp_barrier();
}
template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    simt::check_full(mask);
    (void)width;
    const uint64_t *x = simt::exchange(simt::to_bits(v), mask);
    return simt::from_bits<T>(x[src & 31]);
}
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    simt::check_full(mask);
    (void)width;
    int l = simt::lane();
    const uint64_t *x = simt::exchange(simt::to_bits(v), mask);
    return simt::from_bits<T>(x[(l ^ lanemask) & 31]);
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    simt::check_full(mask);
    (void)width;
    int l = simt::lane();
    const uint64_t *x = simt::exchange(simt::to_bits(v), mask);
    int src = l + (int)delta;
    return simt::from_bits<T>(x[src < 32 ? src : l]);
}
template <typename T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    simt::check_full(mask);
    (void)width;
    int l = simt::lane();
    const uint64_t *x = simt::exchange(simt::to_bits(v), mask);
    int src = l - (int)delta;
    return simt::from_bits<T>(x[src >= 0 ? src : l]);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    simt::check_full(mask);
    const uint64_t *x = simt::exchange(pred ? 1 : 0, mask);
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if ((mask >> i) & 1u) r |= (unsigned)(x[i] & 1) << i;
    return r;
}
template <typename T>
inline unsigned __match_any_sync(unsigned mask, T v) {
    simt::check_full(mask);
    const uint64_t mine = simt::to_bits(v);
    const uint64_t *x = simt::exchange(mine, mask);
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if (((mask >> i) & 1u) && x[i] == mine) r |= 1u << i;
    return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
inline int __reduce_add_sync(unsigned mask, int v) {
    simt::check_full(mask);
    const uint64_t *x = simt::exchange(simt::to_bits(v), mask);
    int r = 0;
    for (int i = 0; i < 32; i++)
        if ((mask >> i) & 1u) r += simt::from_bits<int>(x[i]);
    return r;
}
inline int __reduce_min_sync(unsigned mask, int v) {
    simt::check_full(mask);
    const uint64_t *x = simt::exchange(simt::to_bits(v), mask);
    int r = 0x7FFFFFFF;
    for (int i = 0; i < 32; i++)
        if ((mask >> i) & 1u) r = std::min(r, simt::from_bits<int>(x[i]));
    return r;
}
inline int __reduce_max_sync(unsigned mask, int v) {
    simt::check_full(mask);
    const uint64_t *x = simt::exchange(simt::to_bits(v), mask);
    int r = -0x7FFFFFFF - 1;
    for (int i = 0; i < 32; i++)
        if ((mask >> i) & 1u) r = std::max(r, simt::from_bits<int>(x[i]));
    return r;
}

inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
    simt::check_full(mask);
    const uint64_t *x = simt::exchange(simt::to_bits(v), mask);
    unsigned r = 0xFFFFFFFFu;
    for (int i = 0; i < 32; i++)
        if ((mask >> i) & 1u) r = std::min(r, simt::from_bits<unsigned>(x[i]));
    return r;
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    simt::check_full(mask);
    const uint64_t *x = simt::exchange(simt::to_bits(v), mask);
    unsigned r = 0u;
    for (int i = 0; i < 32; i++)
        if ((mask >> i) & 1u) r = std::max(r, simt::from_bits<unsigned>(x[i]));
    return r;
}

template <typename T>
inline T __ldg(const T *p) { return *p; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)(v >> (shift & 31));
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)((v << (shift & 31)) >> 32);
}
namespace simt {
// PTX prmt.b32, default mode: selector nibble bit 3 replicates the sign of the selected byte
inline unsigned prmt(unsigned x, unsigned y, unsigned s) {
    uint64_t v = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0xF;
        unsigned byte = (unsigned)((v >> (8 * (sel & 7))) & 0xFF);
        if (sel & 8) byte = (byte & 0x80) ? 0xFFu : 0u;
        r |= byte << (8 * i);
    }
    return r;
}
}  // namespace simt
// the CUDA intrinsic uses only three bits of every selector nibble (measured on the B200: no sign mode)
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) { return simt::prmt(x, y, s & 0x7777u); }
inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned shift) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return shift >= 32 ? hi : (unsigned)(v >> shift);
}
inline unsigned __vmaxu4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= std::max((a >> (8 * i)) & 0xFFu, (b >> (8 * i)) & 0xFFu) << (8 * i);
    return r;
}
inline unsigned __vmaxu2(unsigned a, unsigned b) {
    return std::max(a & 0xFFFFu, b & 0xFFFFu) | (std::max(a >> 16, b >> 16) << 16);
}
inline unsigned __vimax3_u16x2(unsigned a, unsigned b, unsigned c) { return __vmaxu2(__vmaxu2(a, b), c); }
inline unsigned __vcmpgeu4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int i = 0; i < 4; i++)
        if (((a >> (8 * i)) & 0xFFu) >= ((b >> (8 * i)) & 0xFFu)) r |= 0xFFu << (8 * i);
    return r;
}
using std::max;
using std::min;

template <typename T>
inline T atomicAdd(T *p, T v) { T old = *p; *p = old + v; return old; }
template <typename T>
inline T atomicMax(T *p, T v) { T old = *p; if (v > old) *p = v; return old; }
template <typename T>
inline T atomicXor(T *p, T v) { T old = *p; *p = old ^ v; return old; }
template <typename T>
inline T atomicAnd(T *p, T v) { T old = *p; *p = old & v; return old; }
template <typename T>
inline T atomicOr(T *p, T v) { T old = *p; *p = old | v; return old; }
template <typename T>
inline T atomicCAS(T *p, T cmp, T v) { T old = *p; if (old == cmp) *p = v; return old; }
template <typename T>
inline T atomicExch(T *p, T v) { T old = *p; *p = v; return old; }
inline void __threadfence() {}
inline void __threadfence_block() {}

// ---- the slice of the CUDA runtime the host side uses -----------------------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
constexpr cudaError_t cudaSuccess = 0;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
struct cudaDeviceProp { int major, minor, multiProcessorCount; char name[64]; size_t sharedMemPerBlockOptin; };
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = aligned_alloc(256, (n + 255) & ~(size_t)255); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void *p) { free(p); return 0; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return 0; }
constexpr unsigned cudaStreamNonBlocking = 1;
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
typedef void *cudaEvent_t;
constexpr unsigned cudaEventDisableTiming = 2;
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (void *)1; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (void *)1; return 0; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = malloc(n); return *p ? 0 : 2; }
inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
constexpr unsigned cudaHostAllocDefault = 0, cudaHostAllocWriteCombined = 4;
inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { *p = malloc(n); return *p ? 0 : 2; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return 0; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof *p);
    p->major = 10;
    p->minor = 0;
    p->multiProcessorCount = 4;
    p->sharedMemPerBlockOptin = 227 * 1024;
    strcpy(p->name, "simt_check");
    return 0;
}
inline const char *cudaGetErrorString(cudaError_t) { return "simt_check"; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
