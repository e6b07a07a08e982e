// Microbenchmark of the per-stage synchronisation primitives (single SM): cycles per operation.
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
__global__ void probe(long long* out) {
    __shared__ unsigned long long bars[64];
    const int lane = threadIdx.x;
    if (lane == 0) { for (int i = 0; i < 64; ++i) mbar_init(s32(&bars[i]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    // warm the code paths once
    if (lane == 0) { arrive(s32(&bars[60])); commit(s32(&bars[61])); }
    wait(s32(&bars[60]), 0); wait(s32(&bars[61]), 0);
    __syncwarp();
    long long t0, t1;
    // (a) 16 barriers completed by plain arrives, then waited by ONE lane
    if (lane == 0) for (int i = 0; i < 16; ++i) arrive(s32(&bars[i]));
    __syncwarp();
    t0 = clock64();
    if (lane == 0) for (int i = 0; i < 16; ++i) wait(s32(&bars[i]), 0);
    t1 = clock64(); if (lane == 0) out[0] = (t1 - t0) / 16;
    __syncwarp();
    // (b) 16 completed barriers waited by ALL 32 lanes
    if (lane == 0) for (int i = 16; i < 32; ++i) arrive(s32(&bars[i]));
    __syncwarp();
    t0 = clock64();
    for (int i = 16; i < 32; ++i) wait(s32(&bars[i]), 0);
    t1 = clock64(); if (lane == 0) out[1] = (t1 - t0) / 16;
    __syncwarp();
    // (c) tcgen05.commit (nothing pending) issue cost, then time until the barrier flips
    t0 = clock64();
    if (lane == 0) for (int i = 32; i < 48; ++i) commit(s32(&bars[i]));
    t1 = clock64(); if (lane == 0) out[2] = (t1 - t0) / 16;
    if (lane == 0) { wait(s32(&bars[47]), 0); out[3] = clock64() - t0; }
    __syncwarp();
    // (d) commit + immediate wait round trip, one at a time
    t0 = clock64();
    if (lane == 0) for (int i = 48; i < 56; ++i) { commit(s32(&bars[i])); wait(s32(&bars[i]), 0); }
    t1 = clock64(); if (lane == 0) out[4] = (t1 - t0) / 8;
    // (e) arrive + wait round trip (same thread)
    t0 = clock64();
    if (lane == 0) for (int i = 56; i < 60; ++i) { arrive(s32(&bars[i])); wait(s32(&bars[i]), 0); }
    t1 = clock64(); if (lane == 0) out[5] = (t1 - t0) / 4;
    // (f) tcgen05 fences + __syncwarp
    t0 = clock64();
    for (int i = 0; i < 16; ++i) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); __syncwarp(); }
    t1 = clock64(); if (lane == 0) out[6] = (t1 - t0) / 16;
}
int main() {
    long long* out; long long h[8];
    cudaMalloc(&out, 64);
    for (int r = 0; r < 3; ++r) probe<<<1, 32>>>(out);
    cudaMemcpy(h, out, 56, cudaMemcpyDeviceToHost);
    printf("wait(done) 1 lane: %lld cyc | 32 lanes: %lld cyc | commit issue: %lld cyc (16 commits visible after %lld) | commit+wait: %lld | arrive+wait: %lld | fence+syncwarp: %lld\n",
           h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
