// Monte-Carlo combine + ELBO head + uncertainty outputs, fused with the ONE exchange of the forward path.
//
// What sits directly above the Bayesian layers:
//   main_bayesian.py:46-53   outputs[:,:,j] = log_softmax(net(x)); log_outputs = logmeanexp_j; kl = mean_j kl_j
//   utils.py:14-22           logmeanexp
//   metrics.py:12-14,23-24   ELBO = nll_loss(log_outputs, y, mean) * train_size + beta * kl ; acc
//   uncertainty_estimation.py:70-96   p_hat = softmax (or softplus-normalised, :73-77); pred = mean_t logits;
//                            epistemic = diag((p_hat - p_bar)^T (p_hat - p_bar)) / T; aleatoric = diag(diag(p_bar) - p_hat^T p_hat / T)
//
// The num_ens samples are sharded over ranks (one process per GPU).  Every rank reduces ITS samples to per-(image,
// class) partials, pushes them straight into every peer's receive buffer over NVLink (plain st.global on peer-mapped
// memory obtained through CUDA IPC), raises a per-CTA flag with release semantics, waits for the peers' flags, and
// finishes the reduction + head locally -- one kernel, no NCCL, no host round trip.  The partials are the exact
// (max, sum-exp) pairs of logmeanexp, so the result equals the reference's logmeanexp even where every sample's
// probability underflows (a plain sum of softmaxes does not).
//
// Receive buffer of one rank (all ranks use the same layout):
//   [0, 4096)                       reserved
//   [4096, ...)                     u64 data[2 slots][world][n_planes * B * C + 2]
// Every float travels as ONE 8-byte store {value bits, sequence number} (the "LL" idea of NCCL's low-latency protocol):
// 8-byte stores are single-copy atomic over NVLink, so the receiver simply polls each word until its tag equals the
// launch's sequence number -- no fence, no separate flag, no CTA barrier between the push and the finish.  (Measured
// with a fence.sys + st.release.sys flag per CTA instead: 3.0 + 3.4 us of fences per step on the critical path.)
// Tags make the buffer reusable without a reset: launch k of a rank uses slot k & 1; a peer can be at most one launch
// ahead (it cannot finish launch k+1 before it has OUR launch k+1 words, which we send after we finished reading k).
#pragma once
#include "common.cuh"

namespace bbb {

constexpr int MCX_MAX_RANKS = 16, MCX_MAX_CTAS = 64, MCX_CTRL_BYTES = 4096, MCX_THREADS = 256, MCX_MAX_SLOCAL = 256;

struct McxArgs {
    const float* logits;          // [S_local, B, C] this rank's samples
    const float* kl;              // n_kl device floats whose SUM is the KL of ONE sample (e.g. the per-layer terms; identical for
    int n_kl;                     // every sample); nullable
    unsigned long long* noise_base; unsigned long long noise_inc;   // optional: *noise_base += noise_inc when the launch is done
    int S_local, S_total, B, C;
    int want_moments, normalized; // moments: also exchange sum p, sum p^2, sum logits;  normalized: p_hat = softplus/sum softplus
    const long long* labels;      // [B] int64, nullable
    float train_size, beta;
    int rank, world;
    unsigned char* peer[MCX_MAX_RANKS];   // receive buffers; peer[rank] is the local one
    unsigned int* seq;            // local device counter: launches completed so far
    unsigned int* done;           // local device counter (zeroed once): CTAs finished in this launch
    unsigned int* timeouts;       // local device counter: waits that gave up (a peer never delivered) -- results are then invalid
    unsigned long long timeout_ns;
    double* head_partials;        // [MCX_MAX_CTAS][2] nll sum, correct count
    // outputs (local)
    float* log_outputs;           // [B, C]
    float* kl_out;                // scalar: sum_j kl_j / S_total
    float* pred; float* epistemic; float* aleatoric; float* entropy;   // [B,C] x3, [B]; nullable (need want_moments)
    float* head;                  // [4]: loss, nll, accuracy, beta*kl; nullable (needs labels)
    long long* tl;                // debug timeline slot (nullptr in production)
    long long* trace;             // debug: [CTA][8] %globaltimer stamps of the handshake (nullptr in production)
};

__host__ __device__ inline int mcx_planes(int want_moments) { return want_moments ? 5 : 2; }
__host__ __device__ inline size_t mcx_rank_floats(int B, int C, int want_moments) { return (size_t)mcx_planes(want_moments) * B * C + 2; }
__host__ inline size_t mcx_buffer_bytes(int B, int C, int want_moments, int world) {
    return MCX_CTRL_BYTES + 2 * (size_t)world * mcx_rank_floats(B, C, want_moments) * sizeof(unsigned long long);
}

__device__ __forceinline__ void st_ll(unsigned long long* p, float v, unsigned int seq) {
    const unsigned long long w = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ld_ll(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float softplus_f(float v) { return v > 20.0f ? v : log1pf(expf(v)); }   // F.softplus(beta=1, threshold=20)

// <= 51 registers: an exchange CTA has to fit beside a GEMM CTA (320 x ~120 registers) and a prep CTA of the next step
__global__ void __launch_bounds__(MCX_THREADS, 5)
mc_exchange_kernel(const McxArgs p) {
    // per warp: log-sum-exp (or softplus sum) of each local sample's row -- dynamic, 32 * S_local bytes: next to a 193 KB
    // GEMM CTA and a 26 KB prep CTA of the NEXT step (overlap mode) a fixed 8 KB array did not fit on the SM any more
    extern __shared__ float norm_dyn[];
    __shared__ unsigned int seq_sh;
    __shared__ double red[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = MCX_THREADS / 32;
    const int B = p.B, C = p.C, BC = B * C;
    const int rows_per_cta = (B + gridDim.x - 1) / gridDim.x;
    const int b0 = blockIdx.x * rows_per_cta, b1 = min(B, b0 + rows_per_cta);
    tl_enter(p.tl);
    asm volatile("griddepcontrol.wait;" ::: "memory");      // launched with programmatic serialization: the logits come from the predecessor
    tl_dep(p.tl);
    const bool tracer = p.trace && threadIdx.x == (p.rank + 1) % p.world;      // the thread that talks to the next rank
    long long* tr = p.trace + blockIdx.x * 8;
    if (tracer) tr[0] = (long long)globaltimer_ns();
    if (threadIdx.x == 0) seq_sh = *p.seq + 1u;
    __syncthreads();
    const unsigned int seq = seq_sh;
    const size_t rank_floats = mcx_rank_floats(B, C, p.want_moments);
    const size_t slot_off = (size_t)(seq & 1u) * p.world * rank_floats;     // in words, behind the control block
    const float inv_S = 1.0f / (float)p.S_total;

    const unsigned long long* rx = reinterpret_cast<const unsigned long long*>(p.peer[p.rank] + MCX_CTRL_BYTES) + slot_off;
    const bool solo = p.world == 1;     // one rank: the partials never leave the registers (no buffer round trip, no handshake)
    double nll_acc = 0.0, hit_acc = 0.0;
    // per-row state of the finish (4): running argmax, entropy, the label's log-probability
    struct RowFin { float best; int best_c; float ent, lab_lp; long long lab; };
    auto fin_elem = [&](RowFin& rf, size_t e, int c, float M, float tot, float sp, float sp2, float sl) {
        const float lo = M + logf(tot * inv_S);                           // utils.py:14-22
        p.log_outputs[e] = lo;
        if (lo > rf.best) { rf.best = lo; rf.best_c = c; }
        if ((long long)c == rf.lab) rf.lab_lp = lo;
        if (p.want_moments) {
            const float pbar = sp * inv_S, p2 = sp2 * inv_S;
            if (p.pred) p.pred[e] = sl * inv_S;                           // uncertainty_estimation.py:82-83
            if (p.epistemic) p.epistemic[e] = p2 - pbar * pbar;           // :89-91  (E[p^2] - p_bar^2)
            if (p.aleatoric) p.aleatoric[e] = pbar - p2;                  // :94-95  (p_bar - E[p^2])
            rf.ent -= pbar > 0.0f ? pbar * logf(pbar) : 0.0f;             // H[p_bar] (no reference)
        }
    };
    // row reductions: entropy, the label's log-probability, argmax (first maximal class, like torch.argmax on ties)
    auto fin_row = [&](RowFin& rf, int b) {
        float ent = warp_sum(rf.ent), lab_lp = warp_sum(rf.lab_lp), best = rf.best;
        int best_c = rf.best_c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
            if (ob > best || (ob == best && oc < best_c)) { best = ob; best_c = oc; }
        }
        if (lane == 0) {
            if (p.entropy && p.want_moments) p.entropy[b] = ent;
            if (p.labels) { nll_acc -= (double)lab_lp; hit_acc += (best_c == (int)rf.lab) ? 1.0 : 0.0; }
        }
    };

    // a word of the receive buffer, once its tag says it belongs to this launch (never hangs: a lost peer = a counted timeout)
    auto ll_value = [&](unsigned long long v, const unsigned long long* w) -> float {
        if ((unsigned int)(v >> 32) != seq) {
            const unsigned long long t0 = globaltimer_ns();
            unsigned int spins = 0;
            while ((unsigned int)((v = ld_ll(w)) >> 32) != seq) {
                __nanosleep(64);          // the SM is shared with the next step's kernels (overlap mode): do not hammer the LSU
                if ((++spins & 63u) == 0u && globaltimer_ns() - t0 > p.timeout_ns) { atomicAdd(p.timeouts, 1u); break; }
            }
        }
        return __uint_as_float((unsigned int)v);
    };

    // ---- (1) local partials of this CTA's images, pushed to every rank's receive buffer ---------------------
    for (int b = b0 + warp; b < b1; b += nwarp) {
        for (int s = 0; s < p.S_local; ++s) {                   // row normaliser of every local sample
            const float* row = p.logits + ((size_t)s * B + b) * C;
            float r;
            if (p.normalized) {
                float acc = 0.0f;
                for (int c = lane; c < C; c += 32) acc += softplus_f(row[c]);
                r = warp_sum(acc);
            } else {
                float mx = -INFINITY;
                for (int c = lane; c < C; c += 32) mx = fmaxf(mx, row[c]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                float se = 0.0f;
                for (int c = lane; c < C; c += 32) se += expf(row[c] - mx);
                r = mx + logf(warp_sum(se));
            }
            if (lane == 0) norm_dyn[warp * p.S_local + s] = r;
        }
        __syncwarp();
        RowFin rf{-INFINITY, 0x7fffffff, 0.0f, 0.0f, (solo && p.labels) ? p.labels[b] : -1};
        for (int c = lane; c < C; c += 32) {
            float mx = -INFINITY, acc = 0.0f, sp = 0.0f, sp2 = 0.0f, sl = 0.0f;
            for (int s = 0; s < p.S_local; ++s) {
                const float l = p.logits[((size_t)s * B + b) * C + c];
                float pr, lp;
                if (p.normalized) { pr = softplus_f(l) / norm_dyn[warp * p.S_local + s]; lp = logf(pr); }
                else { lp = l - norm_dyn[warp * p.S_local + s]; pr = expf(lp); }            // log_softmax (main_bayesian.py:49)
                if (lp > mx) { acc = acc * expf(mx - lp) + 1.0f; mx = lp; }  // online logsumexp over the samples
                else if (lp > -INFINITY) acc += expf(lp - mx);               // lp == -inf: a probability of exactly 0 adds nothing
                sp += pr; sp2 += pr * pr; sl += l;
            }
            const size_t e = (size_t)b * C + c;
            if (solo) { fin_elem(rf, e, c, mx, acc, sp, sp2, sl); continue; }
            for (int q = 0; q < p.world; ++q) {
                unsigned long long* dst = reinterpret_cast<unsigned long long*>(p.peer[q] + MCX_CTRL_BYTES) + slot_off + (size_t)p.rank * rank_floats;
                st_ll(dst + e, mx, seq); st_ll(dst + BC + e, acc, seq);
                if (p.want_moments) { st_ll(dst + 2 * (size_t)BC + e, sp, seq); st_ll(dst + 3 * (size_t)BC + e, sp2, seq); st_ll(dst + 4 * (size_t)BC + e, sl, seq); }
            }
        }
        if (solo) fin_row(rf, b);
        __syncwarp();
    }
    float kl_solo = 0.0f;
    if (solo) {
        if (blockIdx.x == 0 && threadIdx.x == 0) for (int i = 0; p.kl && i < p.n_kl; ++i) kl_solo += __ldg(p.kl + i);
        kl_solo *= (float)p.S_local;
    } else {
        if (blockIdx.x == 0 && threadIdx.x < p.world) {             // this rank's KL contribution: S_local * kl
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(p.peer[threadIdx.x] + MCX_CTRL_BYTES) + slot_off + (size_t)p.rank * rank_floats;
            float one = 0.0f;
            for (int i = 0; p.kl && i < p.n_kl; ++i) one += __ldg(p.kl + i);
            st_ll(dst + (size_t)mcx_planes(p.want_moments) * BC, (float)p.S_local * one, seq);
        }
        if (tracer) tr[1] = (long long)globaltimer_ns();
        // ---- (4) finish: fixed rank order => bitwise identical on every rank ----------------------------------
        for (int b = b0 + warp; b < b1; b += nwarp) {
            RowFin rf{-INFINITY, 0x7fffffff, 0.0f, 0.0f, p.labels ? p.labels[b] : -1};
            for (int c = lane; c < C; c += 32) {
                const size_t e = (size_t)b * C + c;
                // the words of this element from up to 8 ranks in flight at once (one L2 round trip), stragglers polled;
                // ranks merged in ascending order with the online logsumexp update: same operations on every rank
                float M = -INFINITY, tot = 0.0f, mom[3] = {0.0f, 0.0f, 0.0f};
                for (int q0 = 0; q0 < p.world; q0 += 8) {
                    unsigned long long wm[8], wa[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (q0 + j >= p.world) continue;
                        const unsigned long long* r = rx + (size_t)(q0 + j) * rank_floats + e;
                        wm[j] = ld_ll(r); wa[j] = ld_ll(r + BC);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (q0 + j >= p.world) continue;
                        const unsigned long long* r = rx + (size_t)(q0 + j) * rank_floats + e;
                        const float mq = ll_value(wm[j], r), aq = ll_value(wa[j], r + BC);
                        if (aq > 0.0f) {
                            if (mq > M) { tot = tot * expf(M - mq) + aq; M = mq; }
                            else tot += aq * expf(mq - M);
                        }
                    }
                }
                if (p.want_moments) {
#pragma unroll
                    for (int pl = 0; pl < 3; ++pl)
                        for (int q0 = 0; q0 < p.world; q0 += 8) {
                            unsigned long long w[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (q0 + j < p.world) w[j] = ld_ll(rx + (size_t)(q0 + j) * rank_floats + (size_t)(2 + pl) * BC + e);
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (q0 + j < p.world) mom[pl] += ll_value(w[j], rx + (size_t)(q0 + j) * rank_floats + (size_t)(2 + pl) * BC + e);
                        }
                }
                const float sp = mom[0], sp2 = mom[1], sl = mom[2];
                fin_elem(rf, e, c, M, tot, sp, sp2, sl);
            }
            fin_row(rf, b);
        }
    }
    // ---- (5) cross-CTA finish (deterministic order), KL, ELBO head, sequence number ------------------------
    if (solo && !(p.head && p.labels)) {
        // nothing crosses CTAs: KL / Philox base by one thread; the sequence number (slot choice, handshake) is unused
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (p.kl_out) *p.kl_out = kl_solo * inv_S;                    // main_bayesian.py:51  (kl / num_ens)
            if (p.noise_base) *p.noise_base += p.noise_inc;
        }
        tl_exit(p.tl);
        return;
    }
    if (tracer) tr[5] = (long long)globaltimer_ns();
    const bool want_head = p.head && p.labels;
    double nll_cta = 0.0, hit_cta = 0.0;
    if (want_head) {
        nll_cta = block_sum(nll_acc, red);
        __syncthreads();
        hit_cta = block_sum(hit_acc, red);
    } else {
        __syncthreads();                 // the last CTA's bookkeeping below must follow every thread's reads of this launch's slot
    }
    if (threadIdx.x == 0) {
        // (a fence here also waits for the acknowledgements of this CTA's NVLink stores, ~3 us: only where something is published)
        if (want_head) { p.head_partials[2 * blockIdx.x] = nll_cta; p.head_partials[2 * blockIdx.x + 1] = hit_cta; __threadfence(); }
        const unsigned int prev = atomicAdd(p.done, 1u);
        if (prev == gridDim.x - 1) {
            if (want_head) __threadfence();
            float klsum = 0.0f;
            if (solo) {
                for (int i = 0; p.kl && i < p.n_kl; ++i) klsum += __ldg(p.kl + i);
                klsum *= (float)p.S_local;
            } else {
                for (int q = 0; q < p.world; ++q) { const unsigned long long* w = rx + (size_t)q * rank_floats + (size_t)mcx_planes(p.want_moments) * BC; klsum += ll_value(ld_ll(w), w); }
            }
            const float kl = klsum * inv_S;                               // main_bayesian.py:51  (kl / num_ens)
            if (p.kl_out) *p.kl_out = kl;
            if (p.head && p.labels) {
                double nll = 0.0, hit = 0.0;
                for (unsigned int i = 0; i < gridDim.x; ++i) { nll += ((volatile double*)p.head_partials)[2 * i]; hit += ((volatile double*)p.head_partials)[2 * i + 1]; }
                const float nllf = (float)(nll / (double)B);              // F.nll_loss(..., reduction='mean')
                p.head[0] = nllf * p.train_size + p.beta * kl;            // metrics.py:14
                p.head[1] = nllf;
                p.head[2] = (float)(hit / (double)B);                     // metrics.py:23-24
                p.head[3] = p.beta * kl;
            }
            if (p.noise_base) *p.noise_base += p.noise_inc;               // the next step's kernels draw fresh Philox streams
            *p.done = 0u;
            *p.seq = seq;
        }
    }
    tl_exit(p.tl);
}

}  // namespace bbb
