// Microbenchmark (tuning aid, not part of the product): how fast can one persistent CTA per SM stream a buffer into shared
// memory with cp.async.bulk (UBLKCP), by copy size and stages in flight; and with plain LDG.128 for comparison.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) bulk_kernel(const uint8_t *src, size_t bytes, int copy_bytes, int n_stages, unsigned long long *sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bar = (uint64_t *)smem;
    uint8_t *buf = smem + 128;
    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar + s)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const size_t n_copies = bytes / copy_bytes;
    size_t issued = 0, done = 0;
    unsigned long long acc = 0;
    for (size_t t = blockIdx.x; t < n_copies || done < issued; t += gridDim.x) {
        if (t < n_copies) {
            if (issued >= (size_t)n_stages) {  // wait for the oldest
                const int s = (int)(done % n_stages);
                const uint32_t par = (uint32_t)((done / n_stages) & 1);
                asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(smem_u32(bar + s)), "r"(par) : "memory");
                acc += buf[(size_t)s * copy_bytes];
                done++;
            }
            const int s = (int)(issued % n_stages);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar + s)), "r"(copy_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf + (size_t)s * copy_bytes)),
                         "l"(src + t * copy_bytes), "r"(copy_bytes), "r"(smem_u32(bar + s)) : "memory");
            issued++;
        } else {
            const int s = (int)(done % n_stages);
            const uint32_t par = (uint32_t)((done / n_stages) & 1);
            asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D2;\nbra W2;\nD2:\n}\n" ::"r"(smem_u32(bar + s)), "r"(par) : "memory");
            done++;
        }
    }
    if (acc == 0x1234567) *sink = acc;
}
__global__ void __launch_bounds__(512, 1) ldg_kernel(const uint4 *src, size_t n16, unsigned long long *sink) {
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = src[i];
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x1234567) *sink = acc;
}
int main() {
    const size_t bytes = (size_t)448 << 20;
    uint8_t *src; unsigned long long *sink;
    cudaMalloc(&src, bytes); cudaMalloc(&sink, 8); cudaMemset(src, 1, bytes);
    cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int sizes[] = {4096, 8192, 16384, 32768, 49152};
    for (int cs : sizes)
        for (int st = 1; st <= 8; st++) {
            if ((size_t)cs * st + 128 > 227 * 1024) continue;
            if (st != 1 && st != 2 && st != 4 && st != 6 && st != 8 && (size_t)cs * (st + 1) + 128 <= 227 * 1024) continue;
            float best = 1e9;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                bulk_kernel<<<148, 128, cs * st + 128>>>(src, bytes, cs, st, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("bulk copy %6d B x %d stages: %.3f ms  %.0f GB/s  (%s)\n", cs, st, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    for (int grid : {148, 296, 592}) {
        float best = 1e9;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            ldg_kernel<<<grid, 512>>>((const uint4 *)src, bytes / 16, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("LDG.128 grid %d x 512: %.3f ms  %.0f GB/s\n", grid, best, bytes / best / 1e6);
    }
    return 0;
}
