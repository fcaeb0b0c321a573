// N3 - Plackett-Luce re-ranking for the stochastic online simulation feed (sm_100a).
//
// Replaces the per-query host loop of StochasticOnlineSimulationFeed.simulate_clicks_online
// (ultra/input_layer/stochastic_online_simulation_feed.py:100-177): for every list, `np.random.choice(list_len,
// replace=False, p=softmax(tau * scores))` - a Plackett-Luce sample - on the host copy of the scores.  Sampling without
// replacement with probabilities proportional to exp(tau * s_i) is exactly "sort by tau * s_i + Gumbel noise"
// (Gumbel-top-k), so one CTA per list draws counter-based Philox noise, perturbs the scores and ranks them by counting.
// Parity with the reference is distributional (the reference draws from numpy's global RNG), tested on the
// permutation frequencies.
#include "common.cuh"

namespace ub200 {

// ---- Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3") ---------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}
// uniform in the OPEN interval (0, 1): 24 random bits, centred
__device__ __forceinline__ float u01_open(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// scores [B, L]; docid [L, B] (position-major, PAD id == n_docs) or nullptr (all L positions valid);
// perm [B, L]: perm[b][r] = original position of the document shown at rank r; positions >= list_len stay in place.
__global__ void __launch_bounds__(256) pl_sample_kernel(const float* __restrict__ scores,
                                                         const int32_t* __restrict__ docid, int n_docs, int B, int L,
                                                         float tau, unsigned long long seed, unsigned long long offset,
                                                         int32_t* __restrict__ perm) {
    extern __shared__ float key[];     // [L] perturbed scores of the valid positions
    __shared__ int s_len;
    __shared__ float s_red[8];
    griddep_launch();
    griddep_wait();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) s_len = docid ? 0 : L;
        __syncthreads();
        // list_len = 1 + last position holding a real document (stochastic_online_simulation_feed.py:120-127)
        if (docid) {
            int last = 0;
            for (int l = threadIdx.x; l < L; l += blockDim.x)
                if (docid[(size_t)l * B + b] < n_docs) last = l + 1;
            if (last) atomicMax(&s_len, last);
        }
        __syncthreads();
        const int len = s_len;
        const float* s = scores + (size_t)b * L;
        // max over the valid scores (the reference subtracts it before exp, :131-132)
        float m = -INFINITY;
        for (int l = threadIdx.x; l < len; l += blockDim.x) m = fmaxf(m, s[l]);
        m = warp_max(m);
        if (lane == 0) s_red[wid] = m;
        __syncthreads();
        m = s_red[0];
        for (int q = 1; q < nw; ++q) m = fmaxf(m, s_red[q]);
        for (int l = threadIdx.x; l < len; l += blockDim.x) {
            const uint4 r = philox4x32_10(make_uint4((uint32_t)l, (uint32_t)b, (uint32_t)offset, (uint32_t)(offset >> 32)),
                                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
            const float g = -logf(-logf(u01_open(r.x)));                  // standard Gumbel
            key[l] = tau * (s[l] - m) + g;
        }
        __syncthreads();
        int32_t* out = perm + (size_t)b * L;
        for (int l = threadIdx.x; l < L; l += blockDim.x) {
            if (l < len) {
                const float kl = key[l];
                int r = 0;
                for (int j = 0; j < len; ++j) {
                    const float kj = key[j];
                    r += (kj > kl) || (kj == kl && j < l);
                }
                out[r] = l;
            } else {
                out[l] = l;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// N1 - click simulation + batch assembly on the device (position-biased click model).
//
// Replaces ClickSimulationFeed.get_batch (ultra/input_layer/click_simulation_feed.py:101-174) and
// PositionBiasedModel.sampleClicksForOneList (ultra/utils/click_models.py:80-110) for a data set whose initial
// lists, relevance labels and features are resident in HBM: one warp per batch slot draws a query uniformly,
// simulates a click on every position with P = exam_prob[min(l, last)] * click_prob[label] (PAD positions count as
// label 0, as in the reference), and - with check_validation - repeats until the list has at least one click: the
// accepted (query, clicks) pairs have exactly the distribution of the reference's "skip lists without clicks until the
// batch is full" loop.  Output goes straight into the step's staging buffers (doc ids position-major, labels
// list-major); nothing touches the host.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) click_batch_kernel(const int32_t* __restrict__ init_list,
                                                           const float* __restrict__ rel, int nq, int L,
                                                           const float* __restrict__ exam_prob, int n_exam,
                                                           const float* __restrict__ click_prob, int n_cp,
                                                           int oracle_mode, int check_validation, int max_rounds, int B,
                                                           int pad_id, unsigned long long seed,
                                                           unsigned long long offset, int32_t* __restrict__ docid,
                                                           float* __restrict__ labels, int32_t* __restrict__ query_idx,
                                                           int click_model) {
    griddep_launch();
    griddep_wait();
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    if (click_model != 0 && !oracle_mode) {
        // cascade (1) / user-browsing (2) model: a click depends on the clicks before it (click_models.py:112-236), so
        // lane 0 walks the list and keeps the clicks as a bit mask (L <= 256); the draws are still counter based
        // (position, slot, round), the accepted list is written by all lanes
        const uint2 key2 = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        const uint32_t o_lo = (uint32_t)offset, o_hi = (uint32_t)(offset >> 32);
        int q2 = 0;
        uint32_t bits[8];
        for (int round = 0; round < max_rounds; ++round) {
            const uint4 rq = philox4x32_10(make_uint4(0xFFFFFFFFu, (uint32_t)b, o_lo, (o_hi << 12) ^ (uint32_t)round), key2);
            q2 = (int)(((unsigned long long)rq.x * (unsigned long long)nq) >> 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) bits[i] = 0u;
            if (lane == 0) {
                int last = -1;
                bool clicked = false;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint32_t word = 0u;
                    for (int jj = 0; jj < 32; ++jj) {
                        const int l = 32 * i + jj;
                        if (l >= L) break;
                        const float y = rel[(size_t)q2 * L + l];
                        int yi = y > 0.f ? (int)y : 0;
                        if (yi >= n_cp) yi = n_cp - 1;
                        float exam;
                        if (click_model == 1) {
                            exam = exam_prob[l < n_exam ? l : n_exam - 1];
                        } else {
                            // getExamProb(rank, last_click_rank), click_models.py:174-185; exam_prob = [n_exam x n_exam] table
                            const int distance = l - last;
                            if (l < n_exam) exam = exam_prob[l * n_exam + distance - 1];
                            else if (distance > l) exam = exam_prob[(n_exam - 1) * n_exam + n_exam - 1];
                            else exam = exam_prob[(n_exam - 1) * n_exam + (distance < n_exam - 1 ? distance - 1 : n_exam - 2)];
                        }
                        const uint4 r = philox4x32_10(make_uint4((uint32_t)l, (uint32_t)b, o_lo, (o_hi << 12) ^ (uint32_t)round), key2);
                        const bool c = u01_open(r.x) < exam * click_prob[yi];
                        if (c && !(click_model == 1 && clicked)) word |= 1u << jj;    // cascade: only the first click counts
                        if (c) {
                            last = l;
                            clicked = true;
                        }
                    }
                    bits[i] = word;
                }
            }
            uint32_t any = 0u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                bits[i] = __shfl_sync(0xffffffffu, bits[i], 0);
                any |= bits[i];
            }
            if (!check_validation || any || round == max_rounds - 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int l = 32 * i + lane;
                    if (l < L) {
                        const int id = init_list[(size_t)q2 * L + l];
                        docid[(size_t)l * B + b] = id >= 0 ? id : pad_id;
                        labels[(size_t)b * L + l] = (float)((bits[i] >> lane) & 1u);
                    }
                }
                break;
            }
        }
        if (lane == 0 && query_idx) query_idx[b] = q2;
        return;
    }
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t off_lo = (uint32_t)offset, off_hi = (uint32_t)(offset >> 32);
    int q = 0;
    for (int round = 0; round < max_rounds; ++round) {
        // one query per (slot, round): counter word 0 = 0xFFFFFFFF is never a position
        const uint4 rq = philox4x32_10(make_uint4(0xFFFFFFFFu, (uint32_t)b, off_lo, (off_hi << 12) ^ (uint32_t)round), key);
        q = (int)(((unsigned long long)rq.x * (unsigned long long)nq) >> 32);
        int any = 0;
        if (!oracle_mode) {
            for (int l = lane; l < L; l += 32) {
                const float y = rel[(size_t)q * L + l];
                int yi = y > 0.f ? (int)y : 0;
                if (yi >= n_cp) yi = n_cp - 1;
                const float p = exam_prob[l < n_exam ? l : n_exam - 1] * click_prob[yi];
                const uint4 r = philox4x32_10(make_uint4((uint32_t)l, (uint32_t)b, off_lo, (off_hi << 12) ^ (uint32_t)round), key);
                any |= (u01_open(r.x) < p) ? 1 : 0;
            }
        } else {
            for (int l = lane; l < L; l += 32) any |= rel[(size_t)q * L + l] > 0.f ? 1 : 0;
        }
        any = __any_sync(0xffffffffu, any);
        if (!check_validation || any || round == max_rounds - 1) {
            // accepted: regenerate the same draws (same counters) and write the list
            for (int l = lane; l < L; l += 32) {
                const int id = init_list[(size_t)q * L + l];
                const float y = rel[(size_t)q * L + l];
                float c;
                if (oracle_mode) {
                    c = y;
                } else {
                    int yi = y > 0.f ? (int)y : 0;
                    if (yi >= n_cp) yi = n_cp - 1;
                    const float p = exam_prob[l < n_exam ? l : n_exam - 1] * click_prob[yi];
                    const uint4 r = philox4x32_10(make_uint4((uint32_t)l, (uint32_t)b, off_lo, (off_hi << 12) ^ (uint32_t)round), key);
                    c = (u01_open(r.x) < p) ? 1.f : 0.f;
                }
                docid[(size_t)l * B + b] = id >= 0 ? id : pad_id;
                labels[(size_t)b * L + l] = c;
            }
            break;
        }
    }
    if (lane == 0 && query_idx) query_idx[b] = q;
}

}  // namespace ub200

using namespace ub200;

// ---------------------------------------------------------------------------------------------------------------
// RegressionEM (ultra/learning_algorithm/regression_EM.py:108-190): E-step posteriors, Bernoulli pseudo-labels, the
// pointwise sigmoid cross-entropy and its gradient, and the M-step statistics of the examination propensities, one pass
// ---------------------------------------------------------------------------------------------------------------
// Per element (list b, position l), with gamma = sigmoid(s) and e = propensity[l]:
//   P(E=1,R=0 | C=0) = e (1 - gamma) / (1 - e gamma)        P(E=0,R=1 | C=0) = (1 - e) gamma / (1 - e gamma)
//   p_r1 = c + (1 - c) P(E=0,R=1|C=0)        label = ceil(p_r1 - u), u ~ U[0,1)   (get_bernoulli_sample, :20-34)
//   loss += BCEWithLogits(s, label)           ds = sigmoid(s) - label              (mean over B*L: the caller divides)
//   S_l  += c + (1 - c) P(E=1,R=0|C=0)       (M-step: e_l <- (1 - eta) e_l + eta S_l / B)
// One warp per list, lanes over the positions; per-warp position accumulators in shared memory, block partials and a
// last-block reduction in fixed order (deterministic).  out = [sum loss, B*L, S_0 .. S_{L-1}]: sums over the lists,
// so data-parallel ranks add their buffers before dividing.  u comes from Philox keyed by (seed, offset, b, l), or from
// `uniforms` [B, L] when given (parity tests replay the reference's draws).
__global__ void __launch_bounds__(256) regression_em_kernel(const float* __restrict__ scores,
                                                             const float* __restrict__ clicks, int B, int L,
                                                             const float* __restrict__ prop,
                                                             const float* __restrict__ uniforms,
                                                             unsigned long long seed, unsigned long long offset,
                                                             float* __restrict__ dscores, float* __restrict__ out,
                                                             unsigned int* counter, float* __restrict__ partials) {
    griddep_launch();
    griddep_wait();
    extern __shared__ float sm[];                  // [nw][L] position accumulators
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* acc = sm + (size_t)wid * L;
    for (int l = lane; l < L; l += kWarp) acc[l] = 0.f;
    __syncwarp();
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    float loss = 0.f;
    for (int b = blockIdx.x * nw + wid; b < B; b += gridDim.x * nw) {
        for (int l = lane; l < L; l += kWarp) {
            const size_t i = (size_t)b * L + l;
            const float s = scores[i], c = clicks[i], e = prop[l];
            const float gamma = 1.0f / (1.0f + expf(-s));
            const float den = 1.0f - e * gamma;
            const float p_e1_r0 = e * (1.0f - gamma) / den;
            const float p_e0_r1 = (1.0f - e) * gamma / den;
            const float p_r1 = c + (1.0f - c) * p_e0_r1;
            float u;
            if (uniforms) {
                u = uniforms[i];
            } else {
                const uint4 r = philox4x32_10(make_uint4((uint32_t)l, (uint32_t)b, (uint32_t)offset, (uint32_t)(offset >> 32)), key);
                u = (float)(r.x >> 8) * (1.0f / 16777216.0f);          // [0, 1) like torch.rand
            }
            const float label = ceilf(p_r1 - u);
            loss += fmaxf(s, 0.f) - s * label + log1pf(expf(-fabsf(s)));
            dscores[i] = gamma - label;
            acc[l] += c + (1.0f - c) * p_e1_r0;
        }
    }
    loss = warp_sum(loss);
    __shared__ float red[8];
    if (lane == 0) red[wid] = loss;
    __syncthreads();
    const int width = 2 + L;
    float* mine = partials + (size_t)blockIdx.x * width;
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int q = 0; q < nw; ++q) t += red[q];
        mine[0] = t;
        mine[1] = 0.f;
    }
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        float t = 0.f;
        for (int q = 0; q < nw; ++q) t += sm[(size_t)q * L + l];
        mine[2 + l] = t;
    }
    if (last_block_ticket(counter, gridDim.x)) {
        for (int k = threadIdx.x; k < width; k += blockDim.x) {
            float t = 0.f;
            for (int q = 0; q < (int)gridDim.x; ++q) t += partials[(size_t)q * width + k];
            out[k] = (k == 1) ? (float)B * (float)L : t;
        }
    }
}

// M-step (regression_EM.py:181-183): e_l <- (1 - eta) e_l + eta * S_l / B with B = out[1] / L (summed over the ranks)
__global__ void regem_update_kernel(float* __restrict__ prop, const float* __restrict__ out, int L, float em_step) {
    griddep_launch();
    griddep_wait();
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < L) prop[l] = (1.0f - em_step) * prop[l] + em_step * (out[2 + l] / (out[1] / (float)L));
}

static int click_batch_impl(const int32_t* init_list, const float* rel, int nq, int L, const float* exam_prob, int n_exam,
                            const float* click_prob, int n_cp, int oracle_mode, int check_validation, int max_rounds,
                            int B, int pad_id, unsigned long long seed, unsigned long long offset, int32_t* docid,
                            float* labels, int32_t* query_idx, void* stream, int click_model) {
    UB_CHECK(init_list && rel && docid && labels && nq > 0 && L > 0 && B > 0, 2, "click_batch: bad arguments");
    UB_CHECK(click_model >= 0 && click_model <= 2, 1, "click_batch: click_model must be 0 (position biased), 1 (cascade) "
             "or 2 (user browsing), got %d", click_model);
    UB_CHECK(click_model == 0 || oracle_mode || L <= 256, 4, "click_batch: the sequential click models handle lists of up "
             "to 256 positions (got %d)", L);
    UB_CHECK(oracle_mode || (exam_prob && click_prob && n_exam > 0 && n_cp > 0), 2, "click_batch: click model missing");
    UB_CHECK(max_rounds >= 1 && max_rounds < (1 << 12), 1, "click_batch: max_rounds must be in [1, 4096)");
    const int grid = (B + 7) / 8;
    launch_k(click_batch_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), init_list, rel, nq, L, exam_prob,
             n_exam, click_prob, n_cp, oracle_mode, check_validation, max_rounds, B, pad_id, seed, offset, docid, labels,
             query_idx, click_model);
    UB_LAUNCH_CHECK("click_batch_kernel");
    return 0;
}

extern "C" UB200_API int ub200_click_batch(const int32_t* init_list, const float* rel, int nq, int L,
                                           const float* exam_prob, int n_exam, const float* click_prob, int n_cp,
                                           int oracle_mode, int check_validation, int max_rounds, int B, int pad_id,
                                           unsigned long long seed, unsigned long long offset, int32_t* docid,
                                           float* labels, int32_t* query_idx, void* stream) {
    return click_batch_impl(init_list, rel, nq, L, exam_prob, n_exam, click_prob, n_cp, oracle_mode, check_validation,
                            max_rounds, B, pad_id, seed, offset, docid, labels, query_idx, stream, 0);
}

// the same with the click model as a parameter: 0 position biased, 1 cascade (exam_prob [n_exam]), 2 user browsing
// (exam_prob = the [n_exam x n_exam] examination table, row = rank, column = distance to the last click - 1)
extern "C" UB200_API int ub200_click_batch_model(const int32_t* init_list, const float* rel, int nq, int L,
                                                 const float* exam_prob, int n_exam, const float* click_prob, int n_cp,
                                                 int click_model, int oracle_mode, int check_validation, int max_rounds,
                                                 int B, int pad_id, unsigned long long seed, unsigned long long offset,
                                                 int32_t* docid, float* labels, int32_t* query_idx, void* stream) {
    return click_batch_impl(init_list, rel, nq, L, exam_prob, n_exam, click_prob, n_cp, oracle_mode, check_validation,
                            max_rounds, B, pad_id, seed, offset, docid, labels, query_idx, stream, click_model);
}

extern "C" UB200_API int ub200_pl_sample(const float* scores, const int32_t* docid, int n_docs, int B, int L, float tau,
                                         unsigned long long seed, unsigned long long offset, int32_t* perm,
                                         void* stream) {
    UB_CHECK(scores && perm && B > 0 && L > 0, 2, "pl_sample: bad arguments");
    const size_t smem = sizeof(float) * (size_t)L;
    UB_CHECK(smem <= 200 * 1024, 4, "pl_sample: list length %d too large", L);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(pl_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int threads = (L + 31) / 32 * 32;
    if (threads > 256) threads = 256;
    int grid = B < 8 * kNumSMs ? B : 8 * kNumSMs;
    launch_k(pl_sample_kernel, grid, threads, smem, static_cast<cudaStream_t>(stream), scores, docid, n_docs, B, L, tau,
             seed, offset, perm);
    UB_LAUNCH_CHECK("pl_sample_kernel");
    return 0;
}

extern "C" UB200_API int ub200_regression_em(const float* scores, const float* clicks, int B, int L, const float* prop,
                                             const float* uniforms, unsigned long long seed, unsigned long long offset,
                                             float* dscores, float* out, void* workspace, size_t workspace_bytes,
                                             void* stream) {
    UB_CHECK(B > 0 && L > 0, 1, "regression_em: bad B=%d L=%d", B, L);
    UB_CHECK(scores && clicks && prop && dscores && out && workspace, 2, "regression_em: null pointer");
    const int grid = (B + 7) / 8 < 2 * kNumSMs ? (B + 7) / 8 : 2 * kNumSMs;
    const size_t need = 256 + sizeof(float) * (size_t)(2 * kNumSMs) * (2 + L);
    UB_CHECK(workspace_bytes >= need, 3, "regression_em: workspace too small (%zu < %zu)", workspace_bytes, need);
    unsigned int* counter = static_cast<unsigned int*>(workspace);
    float* partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
    const size_t smem = sizeof(float) * 8 * (size_t)L;
    UB_CHECK(smem <= 200 * 1024, 4, "regression_em: list length %d too large", L);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(regression_em_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(regression_em_kernel, grid, 256, smem, static_cast<cudaStream_t>(stream), scores, clicks, B, L, prop, uniforms,
             seed, offset, dscores, out, counter, partials);
    UB_LAUNCH_CHECK("regression_em_kernel");
    return 0;
}

extern "C" UB200_API int ub200_regem_update(float* prop, const float* out, int L, float em_step, void* stream) {
    UB_CHECK(prop && out && L > 0, 2, "regem_update: bad arguments");
    launch_k(regem_update_kernel, (L + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), prop, out, L, em_step);
    UB_LAUNCH_CHECK("regem_update_kernel");
    return 0;
}
