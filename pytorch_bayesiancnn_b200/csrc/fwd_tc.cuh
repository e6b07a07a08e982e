// tcgen05 path (placeholder until the UMMA kernel lands).
#pragma once
#include "common.cuh"
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
};
inline bool tc_supported(const bbb_layer_desc&, const Geom&) { return false; }
inline cudaError_t launch_fwd_tc(const TcArgs&, cudaStream_t, int, int* nl) { *nl = 0; return cudaErrorNotSupported; }
}  // namespace bbb
