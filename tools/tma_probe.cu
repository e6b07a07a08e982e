// Microbenchmark: latency / throughput of cp.async.bulk global->shared on one SM (and on all SMs at once).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tools/tma_probe.cu && ./tma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void wait(uint32_t b, uint32_t ph) {
    asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t n, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(n), "r"(bar) : "memory");
}
// one thread: issue `n` copies of `bytes` (each to its own slot + barrier), then wait them in order
__global__ void probe(const char* src, int bytes, int n, int pieces, long long* out) {
    extern __shared__ __align__(1024) char sm[];
    __shared__ unsigned long long bars[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < n; ++i) mbar_init(s32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const char* my = src + (size_t)blockIdx.x * n * bytes;
        long long t0 = clock64();
        for (int i = 0; i < n; ++i) {
            expect(s32(&bars[i]), bytes);
            for (int p = 0; p < pieces; ++p)
                bulk(s32(sm) + i * bytes + p * (bytes / pieces), my + (size_t)i * bytes + p * (bytes / pieces), bytes / pieces, s32(&bars[i]));
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
        for (int i = 0; i < n; ++i) { wait(s32(&bars[i]), 0); if (blockIdx.x == 0) out[1 + i] = clock64() - t0; }
    }
}
int main() {
    char* src; long long* out; long long h[32];
    size_t total = (size_t)256 << 20;
    cudaMalloc(&src, total); cudaMemset(src, 1, total); cudaMalloc(&out, 32 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int grids[] = {1, 148};
    for (int gi = 0; gi < 2; ++gi)
        for (int bytes : {2048, 8192, 16384, 32768, 49152})
            for (int pieces : {1, 4}) {
                int n = 196608 / bytes; if (n > 8) n = 8;
                for (int rep = 0; rep < 3; ++rep) probe<<<grids[gi], 32, n * bytes, 0>>>(src, bytes, n, pieces, out);   // rep>0: L2 warm
                cudaMemcpy(h, out, 32 * 8, cudaMemcpyDeviceToHost);
                printf("grid %3d  copy %6d B x%d pieces, %d in flight: issue %5lld cyc; done at", grids[gi], bytes, pieces, n, h[0]);
                for (int i = 0; i < n; ++i) printf(" %lld", h[1 + i]);
                printf("\n");
            }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
