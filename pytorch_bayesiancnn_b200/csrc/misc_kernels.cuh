// Stand-alone KL forward/backward, Philox fill and the Monte-Carlo combine.
#pragma once
#include "common.cuh"

namespace bbb {

// kl_loss() without a preceding forward: sigma recomputed from rho.
// HBM-bound: reads 8 B per weight once, float4-vectorised when aligned.
__global__ void __launch_bounds__(256)
kl_forward_kernel(const float* __restrict__ w_mu, const float* __restrict__ w_rho, uint64_t n_w,
                  const float* __restrict__ b_mu, const float* __restrict__ b_rho, uint64_t n_b,
                  float pm, float ps, int conv, double* partials, unsigned int* counter, float* kl_out) {
    __shared__ double red[32];
    double acc = 0.0;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    const bool vec = ((((uintptr_t)w_mu) | ((uintptr_t)w_rho)) & 15u) == 0;
    const uint64_t n4 = vec ? (n_w >> 2) : 0;
    for (uint64_t i = tid; i < n4; i += nth) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(w_mu) + i);
        const float4 r = __ldg(reinterpret_cast<const float4*>(w_rho) + i);
        float s = kl_term(m.x, softplus_sigma(r.x), pm, ps, conv);
        s += kl_term(m.y, softplus_sigma(r.y), pm, ps, conv);
        s += kl_term(m.z, softplus_sigma(r.z), pm, ps, conv);
        s += kl_term(m.w, softplus_sigma(r.w), pm, ps, conv);
        acc += (double)s;
    }
    for (uint64_t i = (n4 << 2) + tid; i < n_w; i += nth)
        acc += (double)kl_term(__ldg(w_mu + i), softplus_sigma(__ldg(w_rho + i)), pm, ps, conv);
    for (uint64_t i = tid; i < n_b; i += nth)
        acc += (double)kl_term(__ldg(b_mu + i), softplus_sigma(__ldg(b_rho + i)), pm, ps, conv);
    const double tot = block_sum(acc, red);
    if (threadIdx.x == 0) kl_publish(tot, blockIdx.x, gridDim.x, partials, counter, kl_out);
}

// d kl / d mu = (mu - pm) / sigma^2 ; d kl / d sigma = 1/sigma - ps^2/sigma^3 - (mu-pm)^2/sigma^3 ;
// d sigma / d rho = sigmoid(rho)   (reference convention).
__global__ void __launch_bounds__(256)
kl_backward_kernel(const float* __restrict__ mu, const float* __restrict__ rho, uint64_t n, float pm, float ps,
                   int conv, const float* __restrict__ grad_kl, float* __restrict__ g_mu, float* __restrict__ g_rho) {
    const float go = __ldg(grad_kl);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float m = mu[i], r = rho[i];
        const float s = softplus_sigma(r), d = m - pm;
        const float sg = 1.0f / (1.0f + expf(-r));
        float dm, ds;
        if (conv == BBB_KL_REFERENCE) {
            const float is = 1.0f / s, is3 = is * is * is;
            dm = d * is * is;
            ds = is - ps * ps * is3 - d * d * is3;
        } else {
            dm = d / (ps * ps);
            ds = -1.0f / s + s / (ps * ps);
        }
        g_mu[i] += go * dm;
        g_rho[i] += go * ds * sg;
    }
}

__global__ void __launch_bounds__(256)
philox_fill_kernel(float* __restrict__ out, uint64_t n, NoiseKey key, uint64_t offset) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = normal1(offset + i, key);
}

__global__ void noise_advance_kernel(unsigned long long* base, unsigned long long inc) { *base += inc; }

// main_bayesian.py:46-53 + utils.py:14-22 (+ uncertainty_estimation.py:70-96 moments).
// One CTA per image; warps compute log-sum-exp per MC sample, then one thread per
// class folds the S samples with an online logmeanexp.
__global__ void __launch_bounds__(128)
mc_combine_kernel(const float* __restrict__ logits, int S, int B, int C, float* __restrict__ log_out,
                  float* __restrict__ moments) {
    extern __shared__ float lse[];            // [S]
    const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int s = wid; s < S; s += nw) {
        const float* row = logits + ((size_t)s * B + b) * C;
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, row[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float se = 0.0f;
        for (int c = lane; c < C; c += 32) se += expf(row[c] - mx);
        se = warp_sum(se);
        if (lane == 0) lse[s] = mx + logf(se);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float mx = -INFINITY, acc = 0.0f, sp = 0.0f, sp2 = 0.0f, sl = 0.0f;
        for (int s = 0; s < S; ++s) {
            const float l = logits[((size_t)s * B + b) * C + c];
            const float v = l - lse[s];          // log_softmax
            if (v > mx) { acc = acc * expf(mx - v) + 1.0f; mx = v; }
            else acc += expf(v - mx);
            const float pr = expf(v);
            sp += pr; sp2 += pr * pr; sl += l;
        }
        log_out[(size_t)b * C + c] = mx + logf(acc / (float)S);
        if (moments) {
            const size_t bc = (size_t)B * C, o = (size_t)b * C + c;
            moments[o] = sp; moments[bc + o] = sp2; moments[2 * bc + o] = sl;
        }
    }
}

}  // namespace bbb
