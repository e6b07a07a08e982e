// backward (placeholder)
#pragma once
#include "common.cuh"
namespace bbb {
struct BwdArgs {
    Geom g;
    const float* x; const float* gy; const float* w_mu; const float* w_rho; const float* b_mu; const float* b_rho;
    const float* act_std; const float* eps_a; const float* eps_b;
    NoiseKey key; const unsigned long long* stream_base;
    float* gx; float* g_w_mu; float* g_w_rho; float* g_b_mu; float* g_b_rho;
    int sample, has_bias, variant;
};
inline cudaError_t launch_bwd_simt(const BwdArgs&, cudaStream_t, int, int* nl) { *nl = 0; return cudaErrorNotSupported; }
}  // namespace bbb
