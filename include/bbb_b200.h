/*
 * bbb_b200.h -- C ABI of the B200-native Bayes-by-Backprop layer engine.
 *
 * The reference (kumar-shridhar/PyTorch-BayesianCNN) has no FFI / plugin layer:
 * its boundary for this path is the Python class surface of layers/ (SURVEY.md
 * 8b).  Each entry point below replaces the body of one reference method; the
 * Python host side (pytorch_bayesiancnn_b200/) keeps the reference's class and
 * argument names and calls these through ctypes (see INTEGRATION.md for the
 * stub a maintainer of the reference would add).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  All pointers are DEVICE
 *     pointers unless the name ends in _host.  `cuda_stream` is a cudaStream_t
 *     (CUstream) handle passed as void*; every call is asynchronous on it and
 *     never synchronises, allocates or takes ownership.
 *   - parameters are fp32, reference layout: W_mu/W_rho [Cout, Cin, kh, kw]
 *     (OIHW) or [out, in]; bias_mu/bias_rho [Cout].  Activations are logical
 *     NCHW, contiguous.
 *   - return 0 on success, a negative BBB_E_* code on error;
 *     bbb_last_error() gives the message (thread-local).
 *   - noise: eps pointers NULL  => in-kernel Philox4x32-10 keyed by
 *     (seed, stream_id, flat element index) -- see bbb_philox_normal_fill for
 *     the exact stream definition; non-NULL => that tensor is used (parity mode,
 *     identical eps to the reference's CPU-generator draws).
 */
#ifndef BBB_B200_H_
#define BBB_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BBB_ABI_VERSION 2

enum { BBB_VARIANT_BBB = 0,   /* weight-space sampling   (layers/BBB/...)      */
       BBB_VARIANT_LRT = 1 }; /* local reparameterisation (layers/BBB_LRT/...) */
enum { BBB_DTYPE_F32 = 0, BBB_DTYPE_BF16 = 1 };
enum { BBB_MATH_FP32 = 0,      /* CUDA-core FFMA, IEEE fp32 accumulate          */
       BBB_MATH_BF16_TC = 1,   /* tcgen05 bf16 x bf16 -> fp32 in TMEM           */
       BBB_MATH_AUTO = 2,      /* engine picks per layer shape                  */
       BBB_MATH_TF32_TC = 3 }; /* tcgen05 tf32 x tf32 -> fp32 in TMEM (operands rounded to tf32, 10-bit mantissa: what the
                                  reference's own GPU conv computes by default); per-layer calls only */
enum { BBB_KL_REFERENCE = 0,   /* as executed by the reference: KL(prior||post) */
       BBB_KL_TEXTBOOK = 1 };  /* KL(q||p)                                      */
enum { BBB_ACT_NONE = 0, BBB_ACT_SOFTPLUS = 1, BBB_ACT_RELU = 2 };

enum { BBB_OK = 0, BBB_E_INVALID = -1, BBB_E_UNSUPPORTED = -2, BBB_E_WORKSPACE = -3,
       BBB_E_CUDA = -4 };

/* Geometry + options of one Bayesian layer call.  A linear layer is the
 * degenerate conv: in_h = in_w = kernel = stride = dil = 1, pad = 0,
 * in_channels = in_features, out_channels = out_features, batch = rows. */
typedef struct bbb_layer_desc {
    int32_t batch;
    int32_t in_channels, in_h, in_w;
    int32_t out_channels;
    int32_t kernel_h, kernel_w;
    int32_t stride_h, stride_w;
    int32_t pad_h, pad_w;
    int32_t dil_h, dil_w;
    int32_t variant;        /* BBB_VARIANT_*                                          */
    int32_t sample;         /* 1: stochastic (self.training or sample); 0: mean only */
    int32_t has_bias;
    int32_t act_dtype;      /* BBB_DTYPE_*: dtype of x and y                          */
    int32_t math;           /* BBB_MATH_*                                             */
    int32_t kl_convention;  /* BBB_KL_*                                               */
    int32_t epilogue_act;   /* BBB_ACT_*: activation fused after the layer (0 = none) */
    int32_t pool_k, pool_s; /* max-pool fused after the activation (0 = none)         */
    int32_t reserved[4];
    float prior_mu, prior_sigma;
} bbb_layer_desc;

/* Bytes of caller-allocated scratch a forward/KL call on `desc` needs.  The
 * scratch must be zero-filled ONCE when allocated; calls leave it zeroed where
 * that matters (self-resetting counters). */
size_t bbb_workspace_bytes(const bbb_layer_desc* desc);

/* Replaces BBBConv2d.forward + .kl_loss:
 *   layers/BBB/BBBConv.py:61-83, layers/BBB_LRT/BBBConv.py:62-87.
 * y      : [batch, out_channels, OH, OW]
 * kl_out : nullable; the layer's KL scalar (weights + bias) is WRITTEN here.
 * act_std: nullable, LRT only, fp32 shape of y: sqrt(act_var) (BBB_LRT/BBBConv.py:75)
 *          saved for the backward.
 * eps_a  : BBB: W_eps [Cout,Cin,kh,kw]; LRT: activation eps, shape of y.  NULL => Philox.
 * eps_b  : BBB: bias_eps [Cout]; LRT: unused.                             NULL => Philox.
 * Philox element index: BBB: flat OIHW index for W, |W| + c for bias;
 *                       LRT: NHWC-flat index of y, ((b*OH*OW + pixel)*Cout + c), i.e. the
 *                       eps tensor is fill(numel).view(B,OH,OW,C).permute(0,3,1,2) -- four
 *                       consecutive channels share one Philox call in every kernel.
 * stream_base: nullable DEVICE pointer; when set the effective Philox stream is
 *          stream_id + *stream_base, read by the kernel at run time -- this is how a
 *          captured CUDA graph draws fresh noise on every replay (bbb_noise_advance). */
int bbb_conv2d_forward(const bbb_layer_desc* desc, const void* x,
                       const float* W_mu, const float* W_rho,
                       const float* bias_mu, const float* bias_rho,
                       void* y, float* kl_out, float* act_std,
                       const float* eps_a, const float* eps_b,
                       uint64_t seed, uint64_t stream_id, const uint64_t* stream_base,
                       void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Replaces BBBLinear.forward + .kl_loss:
 *   layers/BBB/BBBLinear.py:54-76, layers/BBB_LRT/BBBLinear.py:56-79.
 * Same arguments; desc must be the degenerate (1x1) geometry. */
int bbb_linear_forward(const bbb_layer_desc* desc, const void* x,
                       const float* W_mu, const float* W_rho,
                       const float* bias_mu, const float* bias_rho,
                       void* y, float* kl_out, float* act_std,
                       const float* eps_a, const float* eps_b,
                       uint64_t seed, uint64_t stream_id, const uint64_t* stream_base,
                       void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Activation layouts of the fused tcgen05 chain (bbb_layer_forward_fused). */
enum { BBB_LAYOUT_NCHW_F32 = 0,      /* reference layout: [B, C, H, W] fp32                       */
       BBB_LAYOUT_PACKED_BF16 = 1,   /* "tiled packed" bf16: the [B, F] matrix, F = H*W*C, column = (h*W + w)*C + c,
                                        C % 64 == 0, stored as [ceil(B/128)][F/64][128 rows x 128 B] with every 16 KB
                                        block in the K-major SWIZZLE_128B shared-memory image (chunk c of row r at
                                        chunk c ^ (r & 7)); pitch arguments carry F.  When the square is carried too
                                        (LRT consumer) each block is [x | x^2] = 32 KB and x_sq / y_sq = base + 8192 elements */
       BBB_LAYOUT_ROWMAJOR_F32 = 2 };/* [B, OH*OW, Cout] fp32 (logits when OH*OW == 1)            */

/* One Bayesian layer of a fused chain: the layer forward + KL (as bbb_conv2d_forward /
 * bbb_linear_forward) with the model file's activation (desc->epilogue_act) and 2x2/2
 * max-pool (desc->pool_k == 2) fused into the epilogue, reading and writing the packed
 * bf16 inter-layer format so the next layer's operand is a plain 2-D TMA box.  Replaces,
 * per [BBBConv2d|BBBLinear, nn.Softplus|nn.ReLU, nn.MaxPool2d(2,2), FlattenLayer] run of
 * children in ModuleWrapper.forward (layers/misc.py:16-18; BayesianAlexNet.py:34-53).
 *   x, x_sq  : input and (LRT, packed input only) its element-wise square
 *   in_pitch : elements per row of a packed input;  prev_hw: for a linear layer fed by a
 *              flattened HxW map, H*W of that map (reference feature order is c*HW + pix)
 *   y, y_sq  : output and (packed output, nullable) its square for a following LRT layer
 * Forward only (math = BBB_MATH_BF16_TC); eps / Philox / KL semantics as the unfused calls.
 * desc->reserved[0] splits the call so the parameter-only half can run on a side stream, off
 * the activation critical path: BBB_FUSED_PREP_ONLY launches just the weight-prep kernel
 * (sigma, eps, bf16 operand tiles, KL -> kl_out; x/y unused), BBB_FUSED_SKIP_PREP just the GEMM
 * kernel (the caller orders it after the prep, e.g. with an event).  0 = both, in order.
 * desc->reserved[1] > 0 folds Monte-Carlo samples into the batch (LRT + Philox only; what
 * uncertainty_estimation.py:38-41 does by repeating the input): row b of the batch is image b % reserved[1] of sample
 * b / reserved[1], whose noise comes from Philox stream stream_id + (b / reserved[1]) * stride, stride = the uint64
 * in reserved[2] (low) / reserved[3] (high) -- bit-identical to separate calls per sample.  The NCHW input of the
 * first layer then holds reserved[1] images (it is not repeated). */
enum { BBB_FUSED_PREP_ONLY = 1, BBB_FUSED_SKIP_PREP = 2 };
int bbb_layer_forward_fused(const bbb_layer_desc* desc,
                            const void* x, const void* x_sq, int32_t in_layout, int32_t in_pitch, int32_t prev_hw,
                            const float* W_mu, const float* W_rho,
                            const float* bias_mu, const float* bias_rho,
                            void* y, void* y_sq, int32_t out_layout, int32_t out_pitch,
                            float* kl_out, const float* eps_a, const float* eps_b,
                            uint64_t seed, uint64_t stream_id, const uint64_t* stream_base,
                            void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Host-only query (no GPU work, no GPU needed): would bbb_layer_forward_fused accept this layer with these layouts?
 * Returns BBB_OK, or the error code the call would return (bbb_last_error() says why).  The host-side planner
 * (fused.plan) asks before it commits a ModuleWrapper child list to the fused chain. */
int bbb_fused_supported(const bbb_layer_desc* desc, int32_t in_layout, int32_t in_pitch, int32_t prev_hw,
                        int32_t out_layout, int32_t out_pitch);

/* Replaces layer.kl_loss() -> metrics.calculate_kl (metrics.py:27-29 with the call
 * binding of layers/BBB/BBBConv.py:80-82) when no forward preceded it: sigma is
 * recomputed from rho.  n_w = |W|, n_b = |bias| (0 if none). */
int bbb_kl_forward(const float* W_mu, const float* W_rho, uint64_t n_w,
                   const float* bias_mu, const float* bias_rho, uint64_t n_b,
                   float prior_mu, float prior_sigma, int32_t kl_convention,
                   float* kl_out, void* workspace, size_t workspace_bytes, void* cuda_stream);

/* d(kl)/d(mu), d(kl)/d(rho), scaled by *grad_kl (device scalar) and ACCUMULATED
 * into g_mu / g_rho. */
int bbb_kl_backward(const float* mu, const float* rho, uint64_t n,
                    float prior_mu, float prior_sigma, int32_t kl_convention,
                    const float* grad_kl, float* g_mu, float* g_rho, void* cuda_stream);

/* Backward of bbb_conv2d_forward / bbb_linear_forward.
 * Regenerates eps from (seed, stream_id) or reads eps_a/eps_b exactly like the
 * forward.  grad_x nullable.  g_* are ACCUMULATED into (caller zeroes).
 * act_std: LRT only, the tensor the forward saved.                            */
int bbb_conv2d_backward(const bbb_layer_desc* desc, const void* x, const void* grad_y,
                        const float* W_mu, const float* W_rho,
                        const float* bias_mu, const float* bias_rho,
                        const float* act_std,
                        const float* eps_a, const float* eps_b,
                        uint64_t seed, uint64_t stream_id, const uint64_t* stream_base,
                        void* grad_x, float* g_W_mu, float* g_W_rho,
                        float* g_bias_mu, float* g_bias_rho,
                        void* workspace, size_t workspace_bytes, void* cuda_stream);
int bbb_linear_backward(const bbb_layer_desc* desc, const void* x, const void* grad_y,
                        const float* W_mu, const float* W_rho,
                        const float* bias_mu, const float* bias_rho,
                        const float* act_std,
                        const float* eps_a, const float* eps_b,
                        uint64_t seed, uint64_t stream_id, const uint64_t* stream_base,
                        void* grad_x, float* g_W_mu, float* g_W_rho,
                        float* g_bias_mu, float* g_bias_rho,
                        void* workspace, size_t workspace_bytes, void* cuda_stream);

/* The engine's noise stream, exposed so the host side of the boundary can draw
 * exactly what a kernel draws: out[i] = N(0,1) lane ((offset+i)&3) of
 * Philox4x32-10(counter = ((offset+i)>>2, stream_id), key = seed), Box-Muller. */
int bbb_philox_normal_fill(float* out, uint64_t n, uint64_t seed, uint64_t stream_id,
                           uint64_t offset, void* cuda_stream);

/* *base += inc on the device (one tiny kernel; put it at the head of a captured
 * graph so each replay moves every layer to a fresh Philox stream). */
int bbb_noise_advance(uint64_t* base, uint64_t inc, void* cuda_stream);

/* Monte-Carlo combine that sits directly above the path (main_bayesian.py:46-53,
 * utils.py:14-22): logits [S, B, C] fp32 -> log_outputs [B, C] =
 * logmeanexp_s(log_softmax(logits[s])).  Also emits per-sample partials
 * (sum_s softmax, sum_s softmax^2, sum_s logits) [3, B, C] if `moments` != NULL
 * (uncertainty_estimation.py:70-96). */
int bbb_mc_combine(const float* logits, int32_t S, int32_t B, int32_t C,
                   float* log_outputs, float* moments, void* cuda_stream);

/* Monte-Carlo combine + ELBO head + uncertainty outputs, FUSED with the one exchange of the forward path
 * (main_bayesian.py:46-61, utils.py:14-22, metrics.py:12-14,23-24, uncertainty_estimation.py:70-96).
 * The num_ens samples are sharded over `world` ranks (one process per GPU); this rank holds `S_local` of the
 * `S_total` samples' logits [S_local, B, C].  One kernel: per-(image, class) partials of the local samples
 * (the exact (max, sum-exp) pair of logmeanexp, + sum p / sum p^2 / sum logits with BBB_MC_MOMENTS) are stored straight
 * into every rank's receive buffer over NVLink (peer-mapped memory, below) as 8-byte words {value, sequence number};
 * the receiver polls each word until its tag matches (no fence, no flag; a lost peer is a counted time-out, never a
 * hang) and the result is finished locally in fixed rank order (bitwise identical on all ranks):
 *   log_outputs [B,C] = logmeanexp_j log_softmax(logits_j)      kl_out = sum_j kl_j / S_total
 *   pred / epistemic / aleatoric [B,C], entropy [B]             (nullable; need BBB_MC_MOMENTS)
 *   head [4] = {nll*train_size + beta*kl, nll, accuracy, beta*kl}   (nullable; needs labels [B] int64)
 * BBB_MC_NORMALIZED: p_hat = softplus(logits) / sum softplus (uncertainty_estimation.py:73-75) instead of softmax.
 * peer_buffers: HOST array of `world` device pointers, one receive buffer per rank (bbb_mc_buffer_bytes each, zero-
 *   filled once; [rank] is the local one; with world == 1 any device allocation will do);
 * state: local device scratch of bbb_mc_state_bytes(), zero-filled once.  Sequence numbers inside make the buffers
 * reusable call after call (and CUDA-graph replay after replay) with no reset.  Every rank must make the same calls.
 * kl: n_kl device floats whose SUM is one sample's KL (e.g. the per-layer scalars the layer calls wrote: the sum over
 *   layers of ModuleWrapper.forward, layers/misc.py:21-23, happens here); n_kl <= 0 means 1.
 * noise_base (nullable): *noise_base += noise_inc when the launch has finished -- the last kernel of a captured step
 *   moves the Philox stream base for the next replay (replaces a leading bbb_noise_advance launch). */
enum { BBB_MC_MOMENTS = 1, BBB_MC_NORMALIZED = 2 };
size_t bbb_mc_buffer_bytes(int32_t B, int32_t C, int32_t flags, int32_t world);
size_t bbb_mc_state_bytes(void);
int bbb_mc_exchange(const float* logits, int32_t S_local, int32_t S_total, int32_t B, int32_t C, const float* kl,
                    int32_t n_kl, int32_t flags, const int64_t* labels, float train_size, float beta, int32_t rank,
                    int32_t world, void* const* peer_buffers, void* state, float* log_outputs, float* kl_out, float* pred,
                    float* epistemic, float* aleatoric, float* entropy, float* head, uint64_t* noise_base,
                    uint64_t noise_inc, void* cuda_stream);

/* Peer-mapped receive buffers for bbb_mc_exchange (one process per GPU, same node): allocate locally, export a
 * 64-byte CUDA-IPC handle, ship it to the peers by any host channel (the Python side uses torch.distributed),
 * import theirs.  These five calls are the only ones in this library that allocate or synchronise. */
int bbb_comm_alloc(size_t bytes, void** dev_ptr);            /* cudaMalloc + zero fill                 */
int bbb_comm_free(void* dev_ptr);
int bbb_comm_export(void* dev_ptr, void* handle64_host);     /* writes 64 bytes                        */
int bbb_comm_import(const void* handle64_host, void** peer_ptr);
int bbb_comm_unimport(void* peer_ptr);

const char* bbb_last_error(void);
int32_t bbb_abi_version(void);
/* Number of kernels this library has launched since load (all entry points). */
uint64_t bbb_launch_count(void);
/* Tile policy of the fused chain's tap-GEMM layers (bbb_layer_forward_fused), read at launch (= graph capture) time:
 * 0 (default) = 128-column tiles only where the grid still covers most of the SMs (best latency of ONE step);
 * 1 = 128-column tiles wherever Cout allows (fewer operand bytes per MAC; best throughput when several independent
 * steps are in flight and fill the SMs a narrow grid leaves idle -- measured 63.8 -> 59.5 us per BBBAlexNet step
 * with four steps in flight, 92 -> 97 us for a single step).  Returns the previous value. */
int32_t bbb_set_wide_tiles(int32_t prefer_wide);

#ifdef __cplusplus
}
#endif
#endif  /* BBB_B200_H_ */
