// Bayesian layer forward on the 5th-gen tensor cores (BBB_MATH_BF16_TC).
//
// Two kernels per layer call:
//
//  (P) weight_prep_kernel  -- HBM-bound, touches every parameter exactly once:
//      sigma = log1p(exp(rho)); closed-form KL reduced to one scalar; BBB: draws eps
//      (external or Philox) and forms W = mu + eps*sigma, LRT: forms (mu, sigma^2);
//      writes bf16 operand tiles to the workspace ALREADY IN the canonical UMMA
//      K-major core-matrix order, one contiguous 8 KB (BBB) / 16 KB (LRT) block per
//      (n-tile, k-block), so the GEMM kernel stages them with a single bulk-TMA copy.
//      Doing this once per weight instead of once per M-tile CTA removes a
//      (#M-tiles)x redundant softplus/Philox (64x for AlexNet conv2 at B=512).  It
//      depends on the parameters only, so in a captured graph it runs on a side
//      branch, off the activation critical path.
//
//  (G) gemm_tc_kernel  -- implicit-GEMM conv / linear on tcgen05:
//      warps 0-3 : A producers -- gather the im2col rows of x (fp32 NCHW), convert to
//                  bf16 (LRT: also x^2), st.shared into the canonical K-major layout,
//                  fence.proxy.async, mbarrier arrive; afterwards the same warps run
//                  the epilogue (tcgen05.ld -> bias / sqrt(var)*eps -> NCHW store)
//      warp 4    : one elected thread issues tcgen05.mma (M=128, N=64, K=16, bf16 ->
//                  fp32 in TMEM; LRT: second accumulator for the variance path) and
//                  tcgen05.commit to release smem stages / publish the accumulator
//      warp 5    : one elected thread stages the prepared weight tiles with
//                  cp.async.bulk (TMA, mbarrier complete_tx)
//
// Replaces layers/BBB/BBBConv.py:61-83, BBB/BBBLinear.py:54-76,
// BBB_LRT/BBBConv.py:62-87, BBB_LRT/BBBLinear.py:56-79, metrics.py:27-29.
#pragma once
#include <cuda_bf16.h>
#include <cstdlib>
#include "common.cuh"
#include "fwd_simt.cuh"   // apply_act

namespace bbb {

struct TcArgs {
    Geom g;
    const void* x; const float* w_mu; const float* w_rho; const float* b_mu; const float* b_rho;
    void* y; float* kl_out; float* act_std;
    const float* eps_a; const float* eps_b;
    NoiseKey key; const unsigned long long* stream_base;
    double* kl_partials; unsigned int* kl_counter;
    float prior_mu, prior_sigma;
    int sample, kl_convention, has_bias, act, act_dtype, variant;
    // prepared-operand workspace
    __nv_bfloat16* wtiles;   // [n_tiles][k_blocks][planes][8 KB tile]  (64 rows x 8 K chunks of 16 bytes: 64 bf16 or 32 tf32 of K)
    float* bias_ws;          // [2][Npad]: row 0 = bias (BBB: sampled; LRT: mu), row 1 = LRT sigma_b^2
    int n_tiles, k_blocks, planes;
    int skip_prep, prep_only;
    int tf32;                // operands as tf32 (fp32 storage, 4 elements per 16-byte K chunk, kind::tf32) instead of bf16
    int stage_x;             // stage the tile's input images in shared memory: 0 no, 1 as fp32, 2 as bf16 (half the
                             // footprint: lets two LRT CTAs share an SM; x^2 is then formed from the bf16 value)
    // fused epilogue (first layer of a fused chain): 2x2 max-pool + packed bf16 output
    void* y_sq; int out_mode, out_pitch, pool;     // out_mode: 0 packed bf16 [B,(pix,c)], 2 NCHW fp32 (default)
    long long* trace;                              // debug: per-CTA clock64 checkpoints (nullptr in production)
    long long* tl_prep; long long* tl_gemm;        // debug: timeline slots of the two launches (nullptr in production)
};

constexpr int TC_BM = 128, TC_BN = 64, TC_BK = 64;
constexpr int TC_TILE_ELEMS = TC_BN * TC_BK;                 // 4096 bf16 = 8 KB
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;                // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 2;                // 8 KB
constexpr int TC_SMEM_LIMIT = 227 * 1024;

inline int tc_planes(int variant, int sample) { return (variant == BBB_VARIANT_LRT && sample) ? 2 : 1; }
inline size_t tc_stage_bytes(int planes) { return (size_t)planes * (TC_A_BYTES + TC_B_BYTES); }
// A K block is 8 chunks of 16 bytes per row whatever the operand type: 64 bf16 or 32 tf32 elements of K.
__host__ __device__ constexpr int tc_bk(bool tf32) { return tf32 ? TC_BK / 2 : TC_BK; }
inline int tc_kpad(const Geom& g, bool tf32 = false) { return (g.K + tc_bk(tf32) - 1) / tc_bk(tf32) * tc_bk(tf32); }
inline int tc_npad(const Geom& g) { return (g.N + TC_BN - 1) / TC_BN * TC_BN; }
inline size_t tc_fixed_smem(const Geom& g) { return 2048 /*two 1 KB alignment slacks*/ + 1024 /*barriers + bias*/ + (size_t)tc_kpad(g) * 8; }
inline int tc_stages(const Geom& g, int planes) {
    const long avail = (long)TC_SMEM_LIMIT - (long)tc_fixed_smem(g);
    long s = avail / (long)tc_stage_bytes(planes);
    if (s > 4) s = 4;
    return (int)s;
}
inline size_t tc_workspace_bytes(const Geom& g) {
    // 2 planes of operand tiles (bf16: kpad * 2 bytes per row, tf32: kpad32 * 4 -- never less) + bias rows
    return (size_t)tc_npad(g) * tc_kpad(g, true) * 2 /*planes*/ * 4 /*tf32*/ + (size_t)2 * tc_npad(g) * 4;
}
inline bool tc_supported(const bbb_layer_desc& d, const Geom& g) {
    if (d.act_dtype != BBB_DTYPE_F32) return false;
    if (g.M < 1 || g.N < 1) return false;
    if (tc_stages(g, 2) < 2) return false;
    if ((long)tc_npad(g) / TC_BN * (tc_kpad(g, true) / tc_bk(true)) > 1 << 20) return false;
    if (tc_npad(g) / TC_BN > 65535) return false;      // n tiles ride on gridDim.y
    return true;
}

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], tf32 x tf32 -> fp32 (operands: fp32 words, the low 13 mantissa bits ignored)
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may start
// while its predecessor in the stream is still running; it must execute pdl_wait() before touching anything
// the predecessor writes.  pdl_trigger() lets the successor's CTAs start filling idle SMs early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    static int v = -1;
    // on by default: inside the captured chain the next GEMM's CTAs start on idle SMs while the previous GEMM is
    // still in its epilogue, so its prologue (barriers, TMEM, K schedule, LRT noise tile) is done by the time its
    // inputs are: -5 us (LRT) / -10 us (BBB) per BBBAlexNet forward (tools/timeline.py).  BBB_B200_PDL=0 turns it off.
    if (v < 0) { const char* e = getenv("BBB_B200_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
// <<<grid, block, smem, stream>>> with the programmatic-stream-serialization attribute when enabled
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// K-major, SWIZZLE_NONE ("interleave") shared-memory matrix descriptor (sm_100):
//   [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4 (stride between
//   the two 16-byte K chunks of one MMA)  [32,46) stride-dim byte offset >> 4 (stride between
//   8-row core-matrix groups)          [46,48) version = 1            [61,64) layout = 0
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: c_format=F32 [4,6), a/b_format=BF16 [7,10)/[10,13), K-major A and B,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::tf32: a/b_format = TF32 (2)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// round-to-nearest tf32 (10-bit mantissa) kept in an fp32 word: the tensor core would otherwise truncate
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return u;
}
// one 16-byte K chunk of an operand row: 8 bf16 or 4 tf32 values
template <bool TF32>
__device__ __forceinline__ uint4 pack_chunk(const float (&v)[8]) {
    if (TF32) return make_uint4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    const __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    return make_uint4(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b),
                      *reinterpret_cast<const uint32_t*>(&c), *reinterpret_cast<const uint32_t*>(&d));
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

enum { OUT_PACKED_BF16 = 0, OUT_ROWMAJOR_F32 = 1, OUT_NCHW_F32 = 2 };

// "Tiled packed" inter-layer activation: the [B, F] bf16 matrix (F = pixels x channels, F % 64 == 0) is stored as
// [B/128 row tiles][F/64 column blocks][128 rows x 128 B], every 16 KB block already in the K-major SWIZZLE_128B
// smem image -- the consumer stages an A tile with ONE 16 KB cp.async.bulk instead of a 128-row tensor-map box
// (measured: the strided box costs ~900 cycles per stage regardless of bytes, stages or CTA count).
// When a following LRT layer also needs x^2, the two planes of a block are interleaved ([block][x | x^2], 32 KB),
// so the consumer stages both with ONE bulk copy (every cp.async.bulk costs ~200 issue cycles).
// Offset (in elements) of the 8-element chunk holding columns [col, col+8) of row b in plane 0:
__device__ __forceinline__ size_t tiled_chunk_offset(int b, int col, int kb_total, int planes) {
    const int r = b & 127, kb = col >> 6, ch = (col & 63) >> 3;
    return ((size_t)(b >> 7) * kb_total + kb) * (size_t)(planes * 128 * 64) + (size_t)r * 64 + (size_t)((ch ^ (r & 7)) << 3);
}

// Epilogue math of the bf16 path: MUFU-based, a handful of instructions (the exact versions in
// fwd_simt.cuh cost ~100 instructions per value and the epilogue warps run at IPC ~0.25).
// |error| <= ~1e-7 absolute: far inside the bf16 rounding of the values they feed.
__device__ __forceinline__ float fast_act(float v, int act) {
    if (act == BBB_ACT_SOFTPLUS) return fmaxf(v, 0.0f) + __logf(1.0f + __expf(-fabsf(v)));   // == nn.Softplus(1, 20)
    if (act == BBB_ACT_RELU) return fmaxf(v, 0.0f);
    return v;
}
__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// issue-only variant + one wait: several TMEM loads in flight instead of one round trip each
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Apply the fused activation to 16 consecutive output channels [n0, n0+16) of image b at
// output position `pos` (pixel, or pooled window) and store them in the requested layout.
struct StoreCfg { void* y; void* y_sq; int out_mode, out_pitch, N, act; };
__device__ __noinline__ void store_row16(const StoreCfg p, int b, int pos, int n0, float (&v)[16], int ohw_out) {
    const int N = p.N;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fast_act(v[j], p.act);
    if (p.out_mode == OUT_PACKED_BF16 && n0 + 16 <= N) {
        const size_t off = (size_t)b * p.out_pitch + (size_t)pos * N + n0;
        uint4* yo = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + off);
        yo[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        yo[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
        if (p.y_sq) {
            uint4* ys = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y_sq) + off);
            ys[0] = make_uint4(pack_bf16(v[0] * v[0], v[1] * v[1]), pack_bf16(v[2] * v[2], v[3] * v[3]), pack_bf16(v[4] * v[4], v[5] * v[5]), pack_bf16(v[6] * v[6], v[7] * v[7]));
            ys[1] = make_uint4(pack_bf16(v[8] * v[8], v[9] * v[9]), pack_bf16(v[10] * v[10], v[11] * v[11]), pack_bf16(v[12] * v[12], v[13] * v[13]), pack_bf16(v[14] * v[14], v[15] * v[15]));
        }
        return;
    }
#pragma unroll 1
    for (int j = 0; j < 16; ++j) {
        const int n = n0 + j;
        if (n >= N) break;
        if (p.out_mode == OUT_PACKED_BF16) {
            const size_t o = (size_t)b * p.out_pitch + (size_t)pos * N + n;
            reinterpret_cast<__nv_bfloat16*>(p.y)[o] = __float2bfloat16_rn(v[j]);
            if (p.y_sq) reinterpret_cast<__nv_bfloat16*>(p.y_sq)[o] = __float2bfloat16_rn(v[j] * v[j]);
        } else if (p.out_mode == OUT_ROWMAJOR_F32) {
            reinterpret_cast<float*>(p.y)[((size_t)b * ohw_out + pos) * N + n] = v[j];
        } else {
            reinterpret_cast<float*>(p.y)[((size_t)b * N + n) * ohw_out + pos] = v[j];
        }
    }
}

// ------------------------------------------------------------ (P) weight prep
// One CTA per (n-tile, k-block) 64x64 tile (grid-stride).  256 threads: item = (row, 8-wide
// K chunk); consecutive threads take consecutive rows so the 16-byte writes are contiguous.
template <int VARIANT, bool TF32>
__global__ void __launch_bounds__(256)
weight_prep_kernel(const TcArgs p) {
    __shared__ double red[32];
    constexpr bool LRT = VARIANT == BBB_VARIANT_LRT;
    constexpr int CE = TF32 ? 4 : 8, BKE = 8 * CE;                  // elements per 16-byte chunk / per K block
    const Geom& g = p.g;
    const NoiseKey nkey = effective_key(p.key, p.stream_base);
    const bool stoch = p.sample != 0;
    const bool do_kl = p.kl_out != nullptr;
    const int npad = p.n_tiles * TC_BN;
    constexpr int PER_TILE = TC_BN * (TC_BK / 8);                   // (row, 8-wide K chunk) items per tile
    const long n_items = (long)p.n_tiles * p.k_blocks * PER_TILE;
    double kl_acc = 0.0;
    tl_enter(p.tl_prep);
    for (long gi = (long)blockIdx.x * blockDim.x + threadIdx.x; gi < n_items; gi += (long)gridDim.x * blockDim.x) {
        const int tile = (int)(gi / PER_TILE), item = (int)(gi - (long)tile * PER_TILE);
        const int nt = tile / p.k_blocks, kb = tile - nt * p.k_blocks;
        uint8_t* dst = reinterpret_cast<uint8_t*>(p.wtiles) + (size_t)tile * p.planes * TC_B_BYTES;
        const int row = item & (TC_BN - 1), chunk = item >> 6;
        const int n = nt * TC_BN + row, k0 = kb * BKE + chunk * CE;
        float w[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, s2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < CE; ++e) {
            const int k = k0 + e;
            float wv = 0.0f, sv = 0.0f;
            if (n < g.N && k < g.K) {
                const size_t wi = (size_t)n * g.K + k;
                const float mu = __ldg(p.w_mu + wi);
                float sigma = 0.0f;
                if (stoch || do_kl) sigma = softplus_sigma_fast(__ldg(p.w_rho + wi));
                if (LRT) { wv = mu; sv = sigma * sigma; }
                else if (stoch) {
                    const float e_ = p.eps_a ? __ldg(p.eps_a + wi) : normal1(wi, nkey);
                    wv = mu + e_ * sigma;
                } else wv = mu;
                if (do_kl) kl_acc += (double)kl_term_fast(mu, sigma, p.prior_mu, p.prior_sigma, p.kl_convention);
            }
            w[e] = wv; s2[e] = sv;
        }
        // canonical K-major core-matrix order inside the 8 KB tile: chunk*1024 + row*16 bytes
        *reinterpret_cast<uint4*>(dst + chunk * (TC_BN * 16) + row * 16) = pack_chunk<TF32>(w);
        if (p.planes == 2) *reinterpret_cast<uint4*>(dst + TC_B_BYTES + chunk * (TC_BN * 16) + row * 16) = pack_chunk<TF32>(s2);
    }
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < npad; n += gridDim.x * blockDim.x) {   // bias
        float bm = 0.0f, bv = 0.0f;
        if (p.has_bias && n < g.N) {
            const float mu = __ldg(p.b_mu + n);
            const float sigma = (stoch || do_kl) ? softplus_sigma_fast(__ldg(p.b_rho + n)) : 0.0f;
            if (LRT) { bm = mu; bv = sigma * sigma; }
            else if (stoch) {
                const float e_ = p.eps_b ? __ldg(p.eps_b + n) : normal1((uint64_t)g.N * g.K + n, nkey);
                bm = mu + e_ * sigma;
            } else bm = mu;
            if (do_kl) kl_acc += (double)kl_term_fast(mu, sigma, p.prior_mu, p.prior_sigma, p.kl_convention);
        }
        p.bias_ws[n] = bm;
        p.bias_ws[npad + n] = bv;
    }
    if (do_kl) {
        const double tot = block_sum(kl_acc, red);
        if (threadIdx.x == 0) kl_publish(tot, blockIdx.x, gridDim.x, p.kl_partials, p.kl_counter, p.kl_out);
    }
    tl_exit(p.tl_prep);
}

// ----------------------------------------------------------------- (G) GEMM
struct TcSmem {      // barrier block at the start of dynamic smem (after 1024-alignment)
    unsigned long long full[4], empty[4], accum;
    uint32_t tmem_base, pad;
    float bias[64], bvar[64];
};

// Largest number of images a 128-row tile can touch (rows ordered image-major).
__host__ __device__ inline int tc_tile_images(int OHW) {
    if (TC_BM % OHW == 0) return TC_BM / OHW;          // tiles start on image boundaries
    return OHW >= TC_BM ? 2 : (TC_BM + OHW - 1) / OHW + 1;
}

template <int VARIANT, bool TF32>
__global__ void __launch_bounds__(320, 2)
gemm_tc_kernel(const TcArgs p, const int stages) {
    constexpr bool LRT = VARIANT == BBB_VARIANT_LRT;
    constexpr int CE = TF32 ? 4 : 8, BKE = 8 * CE;                  // elements per 16-byte chunk / per K block
    extern __shared__ uint8_t smem_raw[];
    const Geom& g = p.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int planes = p.planes;                       // 2 only for LRT && sample
    const bool two = LRT && planes == 2;

    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    TcSmem* ctl = reinterpret_cast<TcSmem*>(sm);
    int2* ktab = reinterpret_cast<int2*>(sm + 1024);
    const int kpad = p.k_blocks * BKE;
    const uint32_t tiles_off = (1024u + (uint32_t)kpad * 8u + 1023u) & ~1023u;
    const uint32_t stage_bytes = (uint32_t)planes * (TC_A_BYTES + TC_B_BYTES);
    // stage layout: [A (16K)] [A^2 (16K, LRT)] [B planes (8K each)]
    const uint32_t a_off = 0, a2_off = TC_A_BYTES, b_off = (uint32_t)planes * TC_A_BYTES;
    float* xs = reinterpret_cast<float*>(sm + tiles_off + (size_t)stages * stage_bytes);   // staged input images (stage_x)
    __nv_bfloat16* xsh = reinterpret_cast<__nv_bfloat16*>(xs);

    const int n_tile = blockIdx.y, m_tile = blockIdx.x;     // M tiles on x: gridDim.x has no 65535 limit
    const int m0 = m_tile * TC_BM, n0 = n_tile * TC_BN;
    const int chw = g.Cin * g.HW;
    const int img0 = m0 / g.OHW;

    long long* tr = p.trace ? p.trace + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 128 : nullptr;
    if (tr && threadIdx.x == 0) tr[0] = clock64();
    tl_enter(p.tl_gemm);
    pdl_trigger();
    // ---- one-time setup ------------------------------------------------------
    for (int k = threadIdx.x; k < kpad; k += blockDim.x) {
        int2 e;
        if (k < g.K) {
            const int c = k / g.KHW, rs = k - c * g.KHW;
            const int r = rs / g.KW, s = rs - r * g.KW;
            e.x = c * g.HW + r * g.DH * g.W + s * g.DW;
            e.y = ((r * g.DH) << 16) | (s * g.DW);
        } else { e.x = 0; e.y = 0x7fff7fff; }
        ktab[k] = e;
    }
    pdl_wait();                                         // everything below reads/writes tensors other kernels touch
    tl_dep(p.tl_gemm);
    if (p.stage_x) {
        // the tile's input images, loaded once and coalesced; the im2col gather then reads shared memory
        // (LDS latency ~30 cycles) instead of issuing 64 dependent-latency global loads per thread and k-block
        const int last = min(m0 + TC_BM - 1, g.M - 1) / g.OHW;
        const int nflt = (last - img0 + 1) * chw;
        const float* src = reinterpret_cast<const float*>(p.x) + (size_t)img0 * chw;
        const bool vec = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)(nflt * 4)) & 15u) == 0;
        if (p.stage_x == 2) {
            if (vec) {
                for (int i = threadIdx.x; i < (nflt >> 2); i += blockDim.x) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
                    reinterpret_cast<uint2*>(xsh)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
                }
            } else {
                for (int i = threadIdx.x; i < nflt; i += blockDim.x) xsh[i] = __float2bfloat16_rn(__ldg(src + i));
            }
        } else if (vec) {
            for (int i = threadIdx.x; i < (nflt >> 2); i += blockDim.x)
                reinterpret_cast<float4*>(xs)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        } else {
            for (int i = threadIdx.x; i < nflt; i += blockDim.x) xs[i] = __ldg(src + i);
        }
    }
    if (threadIdx.x < 64) {
        const int npad_ = p.n_tiles * TC_BN;
        ctl->bias[threadIdx.x] = p.bias_ws[n0 + threadIdx.x];
        ctl->bvar[threadIdx.x] = p.bias_ws[npad_ + n0 + threadIdx.x];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&ctl->full[s]), 256 + 1);      // 256 A-producer threads + the TMA thread
            mbar_init(smem_u32(&ctl->empty[s]), 1);           // one tcgen05.commit
        }
        mbar_init(smem_u32(&ctl->accum), 1);
        fence_barrier_init();
    }
    const uint32_t tmem_cols = two ? 128u : 64u;
    if (warp == 8) tmem_alloc(smem_u32(&ctl->tmem_base), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = ctl->tmem_base;
    if (tr && threadIdx.x == 0) tr[1] = clock64();

    if (warp < 8) {
        // ================= A producers (then epilogue) =========================
        // 8 warps: thread -> (row = t & 127, half = t >> 7); a half owns 4 of the 8 K-chunks of every k-block in the
        // main loop and 32 of the 64 output columns in the epilogue (warps w and w+4 share TMEM lanes 32*(w&3)..)
        const int t = threadIdx.x & 127, half = threadIdx.x >> 7;   // row of the tile == TMEM lane
        const int m = m0 + t;
        const bool mvalid = m < g.M;
        int ih0 = 0, iw0 = 0; long xb = 0;
        int bimg = 0, pix = 0;
        int pwin = 0;                                  // pooled-window index (pool mode)
        if (mvalid) {
            bimg = m / g.OHW; pix = m - bimg * g.OHW;
            int oh, ow;
            if (p.pool) {                               // rows ordered (image, window, 2x2 position): a quad of lanes == one pool window
                pwin = pix >> 2;
                const int wy = pwin / (g.OW >> 1), wx = pwin - wy * (g.OW >> 1);
                oh = 2 * wy + ((pix >> 1) & 1); ow = 2 * wx + (pix & 1);
                pix = oh * g.OW + ow;
            } else { oh = pix / g.OW; ow = pix - oh * g.OW; }
            ih0 = oh * g.SH - g.PH; iw0 = ow * g.SW - g.PW;
            xb = (long)(p.stage_x ? bimg - img0 : bimg) * chw + (long)ih0 * g.W + iw0;
        }
        const float* __restrict__ xp = p.stage_x ? xs : reinterpret_cast<const float*>(p.x);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
            const int s = kb % stages;
            const uint32_t ph = (uint32_t)(kb / stages) & 1u;
            mbar_wait(smem_u32(&ctl->empty[s]), ph ^ 1u);
            uint8_t* st = sm + tiles_off + (size_t)s * stage_bytes;
#pragma unroll 2
            for (int c8 = half * 4; c8 < half * 4 + 4; ++c8) {
                float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int e = 0; e < CE; ++e) {
                    const int2 kt = ktab[kb * BKE + c8 * CE + e];
                    const int ih = ih0 + (kt.y >> 16), iw = iw0 + (kt.y & 0xffff);
                    float val = 0.0f;
                    if (mvalid && (unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W)
                        val = (p.stage_x == 2) ? __bfloat162float(xsh[xb + kt.x]) : xp[xb + kt.x];
                    v[e] = val;
                }
                *reinterpret_cast<uint4*>(st + a_off + c8 * (TC_BM * 16) + t * 16) = pack_chunk<TF32>(v);
                if (two) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] *= v[e];
                    *reinterpret_cast<uint4*>(st + a2_off + c8 * (TC_BM * 16) + t * 16) = pack_chunk<TF32>(v);
                }
            }
            fence_proxy_async();                        // generic-proxy stores -> visible to the tensor core
            mbar_arrive(smem_u32(&ctl->full[s]));
            if (tr && threadIdx.x == 0 && kb == 0) tr[2] = clock64();
        }
        if (tr && threadIdx.x == 0) tr[3] = clock64();

        // ================= epilogue ============================================
        // (1) LRT noise for this row, 8 columns at a time, drawn while the last MMAs drain
        const bool philox = two && !p.eps_a;
        const NoiseKey nkey = effective_key(p.key, p.stream_base);
        mbar_wait(smem_u32(&ctl->accum), 0u);
        tc_fence_after();
        if (tr && threadIdx.x == 0) tr[5] = clock64();
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const int ohw_out = p.pool ? (g.OHW >> 2) : g.OHW;
        const int opix = p.pool ? pwin : pix;
        const bool writer = mvalid && !(p.pool && (threadIdx.x & 3));
#pragma unroll 1
        for (int c0 = half * 32; c0 < half * 32 + 32; c0 += 8) {
            float am[8], av[8], ez[8];
            tmem_ld8(lane_base + (uint32_t)c0, am);
            if (two) tmem_ld8(lane_base + 64u + (uint32_t)c0, av);
            const int nb = n0 + c0;
            if (philox) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (mvalid && nb + 4 * h < g.N) {
                        const uint64_t o4 = ((uint64_t)bimg * g.OHW + pix) * g.N + nb + 4 * h;   // NHWC-flat element index
                        if ((g.N & 3) == 0) z = normal4(o4 >> 2, nkey);
                        else { z.x = normal1(o4, nkey); z.y = normal1(o4 + 1, nkey); z.z = normal1(o4 + 2, nkey); z.w = normal1(o4 + 3, nkey); }
                    }
                    ez[4 * h] = z.x; ez[4 * h + 1] = z.y; ez[4 * h + 2] = z.z; ez[4 * h + 3] = z.w;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = nb + j;
                float val = -INFINITY;
                if (mvalid && n < g.N) {
                    val = am[j] + ctl->bias[c0 + j];
                    if (two) {
                        const size_t o = ((size_t)bimg * g.N + n) * g.OHW + pix;
                        const float var = 1e-16f + (av[j] + ctl->bvar[c0 + j]);
                        const float sd = p.act_std ? sqrtf(var) : fast_sqrt(var);
                        const float e_ = philox ? ez[j] : __ldg(p.eps_a + o);
                        val = val + sd * e_;
                        if (p.act_std) p.act_std[o] = sd;
                    }
                }
                if (p.pool) {                           // 2x2 max-pool across the lane quad
                    val = fmaxf(val, __shfl_xor_sync(0xffffffffu, val, 1));
                    val = fmaxf(val, __shfl_xor_sync(0xffffffffu, val, 2));
                }
                am[j] = val;
            }
            if (!writer) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) am[j] = fast_act(am[j], p.act);   // writers only; act is monotone: act(max) == max(act)
            if (p.out_mode == OUT_PACKED_BF16) {          // tiled packed (N % 64 == 0 guaranteed by the host)
                const size_t off = tiled_chunk_offset(bimg, opix * g.N + nb, p.out_pitch >> 6, p.y_sq ? 2 : 1);
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + off) =
                    make_uint4(pack_bf16(am[0], am[1]), pack_bf16(am[2], am[3]), pack_bf16(am[4], am[5]), pack_bf16(am[6], am[7]));
                if (p.y_sq)
                    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y_sq) + off) =
                        make_uint4(pack_bf16(am[0] * am[0], am[1] * am[1]), pack_bf16(am[2] * am[2], am[3] * am[3]),
                                   pack_bf16(am[4] * am[4], am[5] * am[5]), pack_bf16(am[6] * am[6], am[7] * am[7]));
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = nb + j;
                    if (n >= g.N) continue;
                    reinterpret_cast<float*>(p.y)[((size_t)bimg * g.N + n) * ohw_out + opix] = am[j];
                }
            }
        }
        if (tr && threadIdx.x == 0) tr[6] = clock64();
        tc_fence_before();
    } else if (warp == 8) {
        // ================= MMA issuer ==========================================
        // one MMA = two 16-byte K chunks per row (K = 16 bf16 or 8 tf32): the byte geometry is the same for both types
        constexpr uint32_t idesc = TF32 ? make_idesc_tf32(TC_BM, TC_BN) : make_idesc_bf16(TC_BM, TC_BN);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
            const int s = kb % stages;
            const uint32_t ph = (uint32_t)(kb / stages) & 1u;
            __syncwarp();
            mbar_wait(smem_u32(&ctl->full[s]), ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t st = base + tiles_off + (uint32_t)s * stage_bytes;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint64_t da = make_smem_desc(st + a_off + j * 2 * (TC_BM * 16), TC_BM * 16, 128);
                    const uint64_t db = make_smem_desc(st + b_off + j * 2 * (TC_BN * 16), TC_BN * 16, 128);
                    if (TF32) umma_tf32(tmem, da, db, idesc, (kb | j) ? 1u : 0u);
                    else umma_bf16(tmem, da, db, idesc, (kb | j) ? 1u : 0u);
                    if (two) {
                        const uint64_t da2 = make_smem_desc(st + a2_off + j * 2 * (TC_BM * 16), TC_BM * 16, 128);
                        const uint64_t db2 = make_smem_desc(st + b_off + TC_B_BYTES + j * 2 * (TC_BN * 16), TC_BN * 16, 128);
                        if (TF32) umma_tf32(tmem + 64u, da2, db2, idesc, (kb | j) ? 1u : 0u);
                        else umma_bf16(tmem + 64u, da2, db2, idesc, (kb | j) ? 1u : 0u);
                    }
                }
                umma_commit(smem_u32(&ctl->empty[s]));            // frees the smem stage when the MMAs retire
                if (kb == p.k_blocks - 1) umma_commit(smem_u32(&ctl->accum));
            }
            __syncwarp();
        }
        tc_fence_before();
    } else {
        // ================= weight-tile TMA ======================================
        // whole warp waits (a blocking try_wait with one active lane is woken ~750 cycles late), lane 0 issues
        {
            const uint32_t bytes = (uint32_t)planes * TC_B_BYTES;
            const uint8_t* src0 = reinterpret_cast<const uint8_t*>(p.wtiles) + (size_t)n_tile * p.k_blocks * planes * TC_B_BYTES;
            for (int kb = 0; kb < p.k_blocks; ++kb) {
                const int s = kb % stages;
                const uint32_t ph = (uint32_t)(kb / stages) & 1u;
                __syncwarp();
                mbar_wait(smem_u32(&ctl->empty[s]), ph ^ 1u);
                if (lane == 0) {
                    const uint32_t bar = smem_u32(&ctl->full[s]);
                    mbar_arrive_expect_tx(bar, bytes);
                    bulk_g2s(base + tiles_off + (uint32_t)s * stage_bytes + b_off, src0 + (size_t)kb * planes * TC_B_BYTES, bytes, bar);
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    tc_fence_after();
    if (warp == 8) tmem_dealloc(tmem, tmem_cols);
    if (tr && threadIdx.x == 256) tr[7] = clock64();
    tl_exit(p.tl_gemm, 256);
}

template <int VARIANT, bool TF32>
inline cudaError_t launch_fwd_tc_t(TcArgs a, cudaStream_t st, int* n_launch) {
    const Geom& g = a.g;
    if (!a.skip_prep) {
        const long items = (long)a.n_tiles * a.k_blocks * TC_BN * 8;
        int grid = (int)((items + 255) / 256);
        if (grid > 2048) grid = 2048;
        // Same shared-memory carve-out as the GEMM kernels: an SM only changes its L1/smem split when idle, so prep
        // CTAs running at the default (small-smem) split kept the first GEMM's CTAs off every SM they touched until
        // their grids drained (tools/timeline.py: first GEMM 8 us after its own prep had finished).
        static const bool carve = [] {
            const char* e = getenv("BBB_B200_PREP_CARVEOUT");
            if (e && e[0] == '0') return false;
            cudaFuncSetAttribute(weight_prep_kernel<VARIANT, TF32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            return true;
        }();
        (void)carve;
        weight_prep_kernel<VARIANT, TF32><<<grid, 256, 0, st>>>(a);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        *n_launch += 1;
        a.kl_out = nullptr;
    }
    if (a.prep_only) return cudaSuccess;
    int stages = tc_stages(g, a.planes);
    // short K loops (AlexNet conv1: 6 k-blocks) gain nothing from a deep ring; two stages let two CTAs share an
    // SM (2 x (2 x 48 KB) for LRT), which hides the gather latency of one CTA behind the other and halves the waves
    if (a.k_blocks <= 8 && stages > 2) stages = 2;
    // exact footprint: base-alignment slack + control/k-table (rounded to 1 KB) + ring (+ staged images)
    const size_t tiles_off = (1024 + (size_t)tc_kpad(g, TF32) * 8 + 1023) / 1024 * 1024;
    size_t smem = 1023 + tiles_off + (size_t)stages * tc_stage_bytes(a.planes);
    const size_t xs_elems = (size_t)tc_tile_images(g.OHW) * g.Cin * g.HW;
    a.stage_x = 0;
    if (xs_elems * 4 <= 32 * 1024 && smem + xs_elems * 4 <= (size_t)TC_SMEM_LIMIT) {
        a.stage_x = 1;
        // two CTAs per SM need 2 * (smem + 1 KB reserved) <= 228 KB: try the half-size bf16 staging when fp32 does not fit
        // (never for tf32 operands: the staged copy would already have lost the bits tf32 keeps)
        const size_t per_sm = 228 * 1024;
        if (!TF32 && 2 * (smem + xs_elems * 4 + 1024) > per_sm && 2 * ((smem + xs_elems * 2 + 127) / 128 * 128 + 1024) <= per_sm) a.stage_x = 2;
        smem += xs_elems * (a.stage_x == 2 ? 2 : 4);
    }
    dim3 grid((g.M + TC_BM - 1) / TC_BM, a.n_tiles);
    cudaFuncSetAttribute(gemm_tc_kernel<VARIANT, TF32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<VARIANT, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = launch_pdl(gemm_tc_kernel<VARIANT, TF32>, grid, dim3(320), smem, st, a, stages);
    if (e != cudaSuccess) return e;
    e = cudaGetLastError();
    if (e == cudaSuccess) *n_launch += 1;
    return e;
}

inline cudaError_t launch_fwd_tc(TcArgs a, cudaStream_t st, int n_sm, int* n_launch) {
    const Geom& g = a.g;
    const bool tf32 = a.tf32 != 0;
    a.planes = tc_planes(a.variant, a.sample);
    a.n_tiles = tc_npad(g) / TC_BN;
    a.k_blocks = tc_kpad(g, tf32) / tc_bk(tf32);
    *n_launch = 0;
    (void)n_sm;
    const bool lrt = a.variant == BBB_VARIANT_LRT;
    if (tf32) return lrt ? launch_fwd_tc_t<BBB_VARIANT_LRT, true>(a, st, n_launch) : launch_fwd_tc_t<BBB_VARIANT_BBB, true>(a, st, n_launch);
    return lrt ? launch_fwd_tc_t<BBB_VARIANT_LRT, false>(a, st, n_launch) : launch_fwd_tc_t<BBB_VARIANT_BBB, false>(a, st, n_launch);
}

}  // namespace bbb
