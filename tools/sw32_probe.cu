// Probe: can a tcgen05 K-major SWIZZLE_32B A-descriptor address OVERLAPPING im2col windows of an image
// staged in shared memory (rows 32 B apart = 4 pixels x 4 channels bf16, arbitrary 32-B-aligned start, arbitrary
// 8-row-group stride)?  That is what a stride-4 conv needs to feed the tensor core straight from the staged
// image, with no per-element gather.  Unknown: whether the swizzle XOR is a function of the ABSOLUTE shared-memory
// address (CuTe writes Swizzle<1,4,3> o smem_ptr o layout), of the offset from the descriptor start, or needs
// the base_offset field.  Each variant stores the image under one hypothesis and checks D = A * B^T exactly.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/sw32_probe tools/sw32_probe.cu && tools/sw32_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../pytorch_bayesiancnn_b200/csrc/fwd_tc.cuh"

using namespace bbb;

constexpr int NIMG = 16, ROWS = 15, WP = 40;            // staged image: [img][row][px][4 ch] bf16, 8 B per pixel
constexpr int ROWB = WP * 8, IMGB = ROWS * ROWB;        // 320 B per row, 4800 B per image
constexpr int IMG_BYTES = NIMG * IMGB + 512;            // + slack for windows that run past the last row

__host__ __device__ inline float xval(int idx) { return (float)((idx * 7 + 3) % 13 - 6); }
__host__ __device__ inline float bval(int n, int k) { return (float)((n * 5 + k * 3) % 7 - 3); }

// mode: 0 absolute-address XOR, base_offset 0 | 1 absolute XOR, base_offset=(start>>7)&7 | 2 XOR relative to the
// descriptor start (only that one window is then consistent) | 3 no swizzle at all in the stored data
__global__ void __launch_bounds__(128)
probe_kernel(float* out, int lr, int kc, int mode, int shift128) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ unsigned long long bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = ((raw + 1023u) & ~1023u) + (shift128 ? 128u : 0u);
    uint8_t* sm = smem_raw + (base - raw);
    uint8_t* img = sm;                                   // IMG_BYTES
    uint8_t* bt = sm + ((IMG_BYTES + 1023) / 1024) * 1024;   // B tile: SWIZZLE_NONE K-major, 2 chunks x 64 rows x 16 B
    const uint32_t start = base + (uint32_t)(lr * ROWB + kc * 32);
    for (int i = threadIdx.x; i < IMG_BYTES / 2; i += blockDim.x) {
        const uint32_t L = (uint32_t)i * 2;              // logical byte offset inside the image buffer
        uint32_t P = L;
        if (mode == 0 || mode == 1) P = L ^ ((((base + L) >> 7) & 1u) << 4);
        else if (mode == 2) { const int rel = (int)(base + L) - (int)start; P = (rel >= 0) ? (uint32_t)((int)L ^ (((rel >> 7) & 1) << 4)) : L; }
        *reinterpret_cast<__nv_bfloat16*>(img + P) = __float2bfloat16_rn(xval(i));
    }
    for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) {
        const int n = i / 16, k = i % 16;
        *reinterpret_cast<__nv_bfloat16*>(bt + (k / 8) * 1024 + n * 16 + (k % 8) * 2) = __float2bfloat16_rn(bval(n, k));
    }
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 64);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        uint64_t da = (uint64_t)((start & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(IMGB >> 4) << 32) | (1ull << 46) | (6ull << 61);
        if (mode == 1) da |= (uint64_t)((start >> 7) & 7u) << 49;
        const uint64_t db = make_smem_desc(smem_u32(bt), 1024, 128);
        umma_bf16(tmem, da, db, make_idesc_bf16(128, 64), 0u);
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0u);
    tc_fence_after();
    const int warp = threadIdx.x >> 5;
    for (int c0 = 0; c0 < 64; c0 += 8) {
        float v[8];
        tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        for (int j = 0; j < 8; ++j) out[threadIdx.x * 64 + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 64);
}

int main() {
    float* d_out; cudaMalloc(&d_out, 128 * 64 * 4);
    std::vector<float> h(128 * 64);
    const size_t smem = 1024 + 128 + ((IMG_BYTES + 1023) / 1024) * 1024 + 2048;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char* names[4] = {"absolute XOR, base_offset=0", "absolute XOR, base_offset=(start>>7)&7", "XOR relative to desc start", "no swizzle in data"};
    for (int mode = 0; mode < 4; ++mode) {
        int bad_cfg = 0; double worst = 0;
        for (int shift = 0; shift < 2; ++shift)
            for (int lr = 0; lr < 11; ++lr)
                for (int kc = 0; kc < 3; ++kc) {
                    probe_kernel<<<1, 128, smem>>>(d_out, lr, kc, mode, shift);
                    if (cudaDeviceSynchronize() != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(cudaGetLastError())); return 1; }
                    cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost);
                    double err = 0;
                    for (int m = 0; m < 128; ++m)
                        for (int n = 0; n < 64; ++n) {
                            const int im = m / 8, ow = m % 8;
                            double ref = 0;
                            for (int k = 0; k < 16; ++k) {
                                const int byte = im * IMGB + lr * ROWB + kc * 32 + ow * 32 + k * 2;   // logical window
                                ref += (double)xval(byte / 2) * bval(n, k);
                            }
                            err = fmax(err, fabs(ref - h[m * 64 + n]));
                        }
                    if (err > 1e-3) ++bad_cfg;
                    worst = fmax(worst, err);
                }
        printf("mode %d (%s): %d / 66 configurations wrong, worst |err| %.1f\n", mode, names[mode], bad_cfg, worst);
    }
    return 0;
}
