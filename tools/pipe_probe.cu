// Reproduce the tap_gemm stage pipeline without MMAs: warp 0 issues 16 KB bulk copies into a 4-stage ring,
// warp 1 (32 lanes) waits on full[s], records the time, and "frees" the stage (arrive or tcgen05.commit).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void wait(uint32_t b, uint32_t ph) {
    asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void commit(uint32_t b) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t n, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(n), "r"(bar) : "memory");
}
// mode bit0: free stages with tcgen05.commit instead of arrive; bit1: 8 extra warps parked on a never-completing barrier;
// bit2: waiter is a single lane
__global__ void probe(const char* src, int bytes, int steps, int mode, long long* out) {
    extern __shared__ __align__(1024) char sm[];
    __shared__ unsigned long long full[4], empty[4], never;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(s32(&full[i]), 1); mbar_init(s32(&empty[i]), 1); }
        mbar_init(s32(&never), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (warp == 0) {
        if (lane == 0)
            for (int it = 0; it < steps; ++it) {
                const int s = it & 3;
                wait(s32(&empty[s]), ((it >> 2) & 1) ^ 1);
                expect(s32(&full[s]), bytes);
                bulk(s32(sm) + s * bytes, src + (size_t)(it % 64) * bytes, bytes, s32(&full[s]));
                out[64 + it] = clock64() - t0;
            }
    } else if (warp == 1) {
        for (int it = 0; it < steps; ++it) {
            const int s = it & 3;
            if (!(mode & 4) || lane == 0) wait(s32(&full[s]), (it >> 2) & 1);
            __syncwarp();
            if (lane == 0) {
                out[it] = clock64() - t0;
                if (mode & 1) commit(s32(&empty[s])); else arrive(s32(&empty[s]));
            }
            __syncwarp();
        }
        if (lane == 0) arrive(s32(&never));
    } else if (mode & 2) {
        wait(s32(&never), 0);
    }
}
int main() {
    char* src; long long* out; long long h[128];
    cudaMalloc(&src, 64 << 20); cudaMemset(src, 1, 64 << 20); cudaMalloc(&out, 128 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int mode = 0; mode < 8; ++mode)
        for (int bytes : {16384, 49152}) {
            for (int rep = 0; rep < 3; ++rep) probe<<<1, 320, 4 * bytes>>>(src, bytes, 24, mode, out);
            cudaMemcpy(h, out, 128 * 8, cudaMemcpyDeviceToHost);
            printf("mode %d (commit=%d parked=%d 1lane=%d) %5d B: full@", mode, mode & 1, (mode >> 1) & 1, (mode >> 2) & 1, bytes);
            for (int i = 0; i < 12; ++i) printf(" %lld", h[i]);
            printf(" | issued@");
            for (int i = 0; i < 8; ++i) printf(" %lld", h[64 + i]);
            printf("\n");
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
