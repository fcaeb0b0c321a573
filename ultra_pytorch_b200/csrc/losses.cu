// K2 / K3 - ranking losses (forward value + gradient w.r.t. the scores) for sm_100a.
//
//  K2 softmax_ce / dla_loss : one warp per ranked list, warp-shuffle reductions, coalesced [B,L] rows.
//  K3 lambdarank / pairdebias: one CTA per ranked list, the L x L pair tile is evaluated on the fly from
//                              shared-memory copies of the list (scores, labels, t+/t-), never materialised.
// All cross-list sums are deterministic: per-block partials + "last block reduces in fixed order".
#include "common.cuh"

namespace ub200 {

constexpr int kLossBlocks = 3 * kNumSMs;       // workspace rows; K3 and the long-list DLA kernels launch at most 2 per SM,
                                               // the short-list DLA kernel (<= 80 registers) 3
constexpr int kLossBlocksWide = 8 * kNumSMs;   // K2 (NA / IPW): enough resident warps to cover the HBM latency

// workspace: [counter (uint, padded to 64 floats)] [partials: kLossBlocks x width floats]
struct LossWs {
    unsigned int* counter;
    float* partials;
};
static size_t loss_ws_bytes(int width) {
    const size_t a = sizeof(float) * (size_t)kLossBlocks * width, b = sizeof(float) * (size_t)kLossBlocksWide * 2;
    return 256 + (a > b ? a : b);
}
static LossWs loss_ws(void* ws) {
    LossWs w;
    w.counter = static_cast<unsigned int*>(ws);
    w.partials = reinterpret_cast<float*>(static_cast<char*>(ws) + 256);
    return w;
}

// deterministic reduction of partials[nblocks][width] into out[width], run by the last block
__device__ __forceinline__ void reduce_partials(const float* __restrict__ partials, int nblocks, int stride,
                                                int width, float* __restrict__ out) {
    // 8 independent accumulators (block b goes to b & 7), combined pairwise at the end: a fixed order, and 8 loads in
    // flight per thread instead of a chain of nblocks L2 round trips.  (Register use matters here: this function is
    // inlined into kernels that sit exactly at an occupancy cliff - pairwise_kernel at 64 registers x 512 threads x 2
    // CTAs - hence 8 accumulators and the explicit minimum-blocks launch bounds of the callers.)
    for (int k = threadIdx.x; k < width; k += blockDim.x) {
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.f;
        int b = 0;
        for (; b + 8 <= nblocks; b += 8) {
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] += partials[(size_t)(b + q) * stride + k];
        }
        for (; b < nblocks; ++b) acc[b & 7] += partials[(size_t)b * stride + k];
        out[k] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    }
}

// ------------------------------------------------------------------------------------------------
// K2: softmax cross-entropy (NA / IPW) and the two DLA losses
// ------------------------------------------------------------------------------------------------
// Per list (base_algorithm.py:322-330): w = (y + 1e-7) * pw ; W = sum w ; d = nan_to_num(w / W) ;
//   l = -sum_l d_l * log_softmax(s)_l * W ;  dl/ds_l = (softmax(s)_l * sum(d) - d_l) * W
struct ListStats {
    float lse;    // log-sum-exp of the logits
    float W;      // sum of weighted labels
    float dsum;   // sum of d (1, or 0 for an empty list)
};

// NA / IPW, lists of up to G * EPL positions: a list is owned by a GROUP of G lanes (8, 16 or 32), EPL elements per
// lane, so a warp reduces 32 / G lists at once (4 lists of 40 positions: 3 shuffle steps per reduction instead of 5 and
// no idle half-warp) and the whole list lives in registers: one global read of scores and labels, one exponential
// per element, the next lists prefetched while the current ones are reduced.  Algorithmic traffic 12 L + 8 bytes per
// list; ncu showed the warp-per-list form issue-bound (325 warp instructions per list, 81 % issue-active at 15 % DRAM
// throughput) - this form needs ~60.
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// exp(x) for x <= ~0 as ONE multiply + ex2.approx.ftz (the non-ftz __expf adds a range test and two scalings per call:
// 5 instructions; results below 2^-126 flush to zero, which the softmax sums do not see)
__device__ __forceinline__ float exp_ftz(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
}

template <int MODE, int G, int EPL>   // MODE 0 = no weights, 1 = IPW table
__global__ void __launch_bounds__(256) softmax_ce_reg_kernel(const float* __restrict__ scores,
                                                              const float* __restrict__ labels, int B, int L,
                                                              const float* __restrict__ table, int table_len,
                                                              float* __restrict__ dscores, float* __restrict__ sums,
                                                              unsigned int* counter, float* __restrict__ partials) {
    griddep_launch();
    griddep_wait();
    constexpr int LPW = 32 / G;                      // lists per warp
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int gl = lane % G, gi = lane / G;          // lane inside the group, group inside the warp
    const int stride = gridDim.x * nw * LPW;
    float tw[EPL];                                   // IPW weight of this lane's positions (loop invariant)
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        const int l = gl + G * e;
        tw[e] = (MODE == 1 && l < L) ? table[min(l, table_len - 1)] : 1.f;
    }
    float sv[EPL], yv[EPL], sn[EPL], yn[EPL];
    auto load = [&](int bb, float* s_, float* y_) {
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const int l = gl + G * e;
            const bool ok = bb < B && l < L;
            s_[e] = ok ? __ldg(scores + (size_t)bb * L + l) : -INFINITY;
            y_[e] = ok ? __ldg(labels + (size_t)bb * L + l) : 0.f;
        }
    };
    int b = (blockIdx.x * nw + wid) * LPW + gi;
    load(b, sv, yv);
    float num = 0.f, den = 0.f;
    // every lane of the warp runs the same number of iterations (shuffles); groups past the end work on empty lists
    const int b_warp0 = (blockIdx.x * nw + wid) * LPW;
    for (int bw = b_warp0; bw < B; bw += stride, b += stride) {
        load(b + stride, sn, yn);                    // in flight while these lists are reduced
        const bool live = b < B;
        float m = -INFINITY;
#pragma unroll
        for (int e = 0; e < EPL; ++e) m = fmaxf(m, sv[e]);
        m = group_max<G>(m);
        float pe[EPL], w[EPL];
        float esum = 0.f, W = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const bool ok = live && gl + G * e < L;
            pe[e] = ok ? exp_ftz(sv[e] - m) : 0.f;                      // ex2.approx: |rel err| < 2^-21, argument <= 0
            esum += pe[e];
            float pw = 1.f;
            if (MODE == 1) pw = (yv[e] > 0.f) ? tw[e] : 0.f;           // ipw_rank.py:116-128
            w[e] = ok ? (yv[e] + 1e-7f) * pw : 0.f;
            W += w[e];
        }
        esum = group_sum<G>(esum);
        W = group_sum<G>(W);
        const float lse = m + logf(esum);
        const float inv_e = 1.f / esum;
        const float inv_W = (W != 0.f) ? 1.f / W : 0.f;               // nan_to_num(w / W): one reciprocal per list
        float ll = 0.f, dsum = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const bool ok = live && gl + G * e < L;
            const float d = w[e] * inv_W;
            w[e] = d;
            if (ok) ll = fmaf(-d, sv[e] - lse, ll);
            dsum += d;
        }
        ll = group_sum<G>(ll);
        dsum = group_sum<G>(dsum);
        if (live) {
#pragma unroll
            for (int e = 0; e < EPL; ++e) {
                const int l = gl + G * e;
                if (l < L) dscores[(size_t)b * L + l] = (pe[e] * inv_e * dsum - w[e]) * W;
            }
            if (gl == 0) {
                num += ll * W;
                den += W;
            }
        }
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            sv[e] = sn[e];
            yv[e] = yn[e];
        }
    }
    num = warp_sum(num);                             // lanes with gl != 0 hold zeros
    den = warp_sum(den);
    // block partials [num, den], then the last block adds all of them in a fixed order
    __shared__ float red[8][2];
    __shared__ float wide[256][2];
    if (lane == 0) {
        red[wid][0] = num;
        red[wid][1] = den;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        float t = 0.f;
        for (int q = 0; q < nw; ++q) t += red[q][threadIdx.x];
        partials[(size_t)blockIdx.x * 2 + threadIdx.x] = t;
    }
    if (last_block_ticket(counter, gridDim.x)) {
        float a0 = 0.f, a1 = 0.f;
        for (int q = threadIdx.x; q < (int)gridDim.x; q += blockDim.x) {
            a0 += partials[(size_t)q * 2];
            a1 += partials[(size_t)q * 2 + 1];
        }
        wide[threadIdx.x][0] = a0;
        wide[threadIdx.x][1] = a1;
        __syncthreads();
        if (threadIdx.x < 2) {
            float t = 0.f;
            for (int q = 0; q < (int)blockDim.x; ++q) t += wide[q][threadIdx.x];
            sums[threadIdx.x] = t;
        }
    }
}

// DLA (dla.py:179-306) in the same lane-group register form: both softmax losses of a list - the ranking loss weighted
// by the inverse propensities  softmax(prop)_0 / softmax(prop)_l  and the examination loss weighted by the inverse
// relevances  softmax(s)_0 / softmax(s)_l - from ONE exponential per element (the relevance ratio is pe_0 / pe_l, the
// partition function cancels), one reciprocal per list and weight set, the DenoisingNet outputs prop_l = ELU(W_l + b),
// their softmax and the propensity ratios in registers per lane position (loop invariant), and the gradient w.r.t. the
// propensity logits accumulated in registers over the lists a lane group owns, then combined in fixed order
// (group -> warp -> block -> last block).  The warp-per-list form it replaces made three passes over a list with an
// expf and a division per element and pass (issue-bound: 325 warp instructions per list).
template <int G, int EPL>
__global__ void __launch_bounds__(256) dla_reg_kernel(const float* __restrict__ scores, const float* __restrict__ clicks,
                                                       int B, int L, const float* __restrict__ prop_w,
                                                       const float* __restrict__ prop_b, float* __restrict__ dscores,
                                                       float* __restrict__ dprop, float* __restrict__ sums,
                                                       unsigned int* counter, float* __restrict__ partials) {
    griddep_launch();
    griddep_wait();
    constexpr int LPW = 32 / G;
    __shared__ float s_prop[256];
    __shared__ float s_acc[8][256];
    __shared__ float s_lse[1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int gl = lane % G, gi = lane / G;
    const int stride = gridDim.x * nw * LPW;
    // DenoisingNet forward (dla.py:32-46) + log-partition of its outputs, once per block
    const float pb = prop_b[0];
    for (int l = threadIdx.x; l < L; l += blockDim.x) s_prop[l] = elu_f(prop_w[l] + pb);
    __syncthreads();
    if (wid == 0) {
        float m = -INFINITY;
        for (int l = lane; l < L; l += kWarp) m = fmaxf(m, s_prop[l]);
        m = warp_max(m);
        float e = 0.f;
        for (int l = lane; l < L; l += kWarp) e += expf(s_prop[l] - m);
        e = warp_sum(e);
        if (lane == 0) s_lse[0] = m + logf(e);
    }
    __syncthreads();
    const float lse_p = s_lse[0];
    const float smp0 = expf(s_prop[0] - lse_p);
    float lp[EPL], smp[EPL], pwt[EPL], gacc[EPL];      // log softmax(prop)_l, softmax(prop)_l, propensity ratio, dL/dprop_l
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        const int l = gl + G * e;
        lp[e] = l < L ? s_prop[l] - lse_p : 0.f;
        smp[e] = l < L ? expf(lp[e]) : 1.f;
        pwt[e] = l < L ? smp0 / smp[e] : 0.f;           // get_normalized_weights, dla.py:296-298
        gacc[e] = 0.f;
    }
    float sv[EPL], yv[EPL], sn[EPL], yn[EPL];
    auto load = [&](int bb, float* s_, float* y_) {
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const int l = gl + G * e;
            const bool ok = bb < B && l < L;
            s_[e] = ok ? __ldg(scores + (size_t)bb * L + l) : -INFINITY;
            y_[e] = ok ? __ldg(clicks + (size_t)bb * L + l) : 0.f;
        }
    };
    int b = (blockIdx.x * nw + wid) * LPW + gi;
    load(b, sv, yv);
    float num = 0.f, den = 0.f, num_e = 0.f, den_e = 0.f;
    const int b_warp0 = (blockIdx.x * nw + wid) * LPW;
    for (int bw = b_warp0; bw < B; bw += stride, b += stride) {
        load(b + stride, sn, yn);
        const bool live = b < B;
        float m = -INFINITY;
#pragma unroll
        for (int e = 0; e < EPL; ++e) m = fmaxf(m, sv[e]);
        m = group_max<G>(m);
        float pe[EPL], w[EPL], we[EPL];
        float esum = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const bool ok = live && gl + G * e < L;
            pe[e] = ok ? exp_ftz(sv[e] - m) : 0.f;
            esum += pe[e];
        }
        esum = group_sum<G>(esum);
        const float s0 = __shfl_sync(0xffffffffu, sv[0], gi * G);       // position 0 lives in the group's first lane
        // one reduction phase for everything that is linear in the weights: W = sum w, A = sum w s, and the same for the
        // examination loss (its targets lp are the log-softmax of the propensity logits).  With d = w / W:
        //   -sum d (s - lse) = lse - A / W,   sum d = 1,   (softmax(s) sum d - d) W = softmax(s) W - w
        float W = 0.f, We = 0.f, A = 0.f, Ae = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const bool ok = live && gl + G * e < L;
            const float yl = yv[e] + 1e-7f;
            w[e] = ok ? yl * pwt[e] : 0.f;
            we[e] = ok ? yl * exp_ftz(s0 - sv[e]) : 0.f;                 // softmax(s)_0 / softmax(s)_l
            W += w[e];
            We += we[e];
            A = fmaf(w[e], ok ? sv[e] : 0.f, A);
            Ae = fmaf(we[e], lp[e], Ae);
        }
        W = group_sum<G>(W);
        We = group_sum<G>(We);
        A = group_sum<G>(A);
        Ae = group_sum<G>(Ae);
        if (live) {
            const float lse = m + __logf(esum);
            const float pW = (W != 0.f) ? __fdividef(W, esum) : 0.f;               // W / Z: softmax(s)_l W = pe_l pW
            const float pWe = (We != 0.f) ? We : 0.f;
#pragma unroll
            for (int e = 0; e < EPL; ++e) {
                const int l = gl + G * e;
                if (l < L) {
                    dscores[(size_t)b * L + l] = fmaf(pe[e], pW, -w[e]);
                    gacc[e] += fmaf(smp[e], pWe, -we[e]);
                }
            }
            if (gl == 0) {
                num += (W != 0.f) ? fmaf(lse, W, -A) : 0.f;               // ll W = (lse - A / W) W
                den += W;
                num_e -= (We != 0.f) ? Ae : 0.f;                         // lle We = -(Ae / We) We
                den_e += We;
            }
        }
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            sv[e] = sn[e];
            yv[e] = yn[e];
        }
    }
    // gradient w.r.t. the propensity logits: groups of a warp in fixed order, then warps, then blocks
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
        float t = gacc[e];
#pragma unroll
        for (int o = G; o < 32; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        const int l = gl + G * e;
        if (gi == 0 && l < L) s_acc[wid][l] = t;
    }
    num = warp_sum(num);
    den = warp_sum(den);
    num_e = warp_sum(num_e);
    den_e = warp_sum(den_e);
    __shared__ float red[8][4];
    if (lane == 0) {
        red[wid][0] = num; red[wid][1] = den; red[wid][2] = num_e; red[wid][3] = den_e;
    }
    __syncthreads();
    const int width = 4 + L;
    float* mine = partials + (size_t)blockIdx.x * width;
    if (threadIdx.x < 4) {
        float t = 0.f;
        for (int q = 0; q < nw; ++q) t += red[q][threadIdx.x];
        mine[threadIdx.x] = t;
    }
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        float t = 0.f;
        for (int q = 0; q < nw; ++q) t += s_acc[q][l];
        mine[4 + l] = t;
    }
    if (last_block_ticket(counter, gridDim.x)) {
        // sums[4], then the exam gradient through ELU and the Linear(L, 1): dprop[l] = g_l ELU'(pre_l),
        // dprop[L] = sum_l dprop[l]   (dla.py:24-48)
        __shared__ float gsum[256];
        reduce_partials(partials, gridDim.x, width, 4, sums);
        float local = 0.f;
        for (int l = threadIdx.x; l < L; l += blockDim.x) {
            float acc[8];                      // as reduce_partials: 8 loads in flight, fixed combination order
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = 0.f;
            int b = 0;
            for (; b + 8 <= (int)gridDim.x; b += 8) {
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] += partials[(size_t)(b + q) * width + 4 + l];
            }
            for (; b < (int)gridDim.x; ++b) acc[b & 7] += partials[(size_t)b * width + 4 + l];
            const float g = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
            const float pre = prop_w[l] + pb;
            const float gp = g * (pre > 0.f ? 1.f : expf(pre));
            dprop[l] = gp;
            local += gp;
        }
        gsum[threadIdx.x] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int q = 0; q < (int)blockDim.x; ++q) t += gsum[q];
            dprop[L] = t;
        }
    }
}

template <int MODE>   // 0 = no weights, 1 = IPW table, 2 = DLA
__global__ void __launch_bounds__(256) softmax_ce_kernel(const float* __restrict__ scores,
                                                          const float* __restrict__ labels, int B, int L,
                                                          const float* __restrict__ table, int table_len,
                                                          const float* __restrict__ prop_w,
                                                          const float* __restrict__ prop_b,
                                                          float* __restrict__ dscores, float* __restrict__ dprop,
                                                          float* __restrict__ sums, unsigned int* counter,
                                                          float* __restrict__ partials) {
    extern __shared__ float sm[];
    griddep_launch();
    griddep_wait();
    // DLA shared: prop[L], sm_p[L] (softmax of prop), then per-warp accumulators acc[nw][L]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* prop = sm;
    float* smp = sm + L;
    float* acc = sm + 2 * L + (size_t)wid * L;
    float lse_p = 0.f;
    if (MODE == 2) {
        const float pb = prop_b[0];
        for (int l = threadIdx.x; l < L; l += blockDim.x) prop[l] = elu_f(prop_w[l] + pb);   // dla.py:32-46
        for (int l = lane; l < L; l += kWarp) acc[l] = 0.f;
        __syncthreads();
        float m = -INFINITY;
        for (int l = lane; l < L; l += kWarp) m = fmaxf(m, prop[l]);
        m = warp_max(m);
        float e = 0.f;
        for (int l = lane; l < L; l += kWarp) e += expf(prop[l] - m);
        e = warp_sum(e);
        lse_p = m + logf(e);
        if (wid == 0)
            for (int l = lane; l < L; l += kWarp) smp[l] = expf(prop[l] - lse_p);
        __syncthreads();
    }
    float num = 0.f, den = 0.f, num_e = 0.f, den_e = 0.f;
    for (int b = blockIdx.x * nw + wid; b < B; b += gridDim.x * nw) {
        const float* s = scores + (size_t)b * L;
        const float* y = labels + (size_t)b * L;
        // log-sum-exp of the scores
        float m = -INFINITY;
        for (int l = lane; l < L; l += kWarp) m = fmaxf(m, s[l]);
        m = warp_max(m);
        float e = 0.f;
        for (int l = lane; l < L; l += kWarp) e += expf(s[l] - m);
        e = warp_sum(e);
        const float lse = m + logf(e);
        // weights
        float W = 0.f, We = 0.f;
        float sm_s0 = 0.f, smp0 = 0.f;
        if (MODE == 2) {
            sm_s0 = expf(s[0] - lse);
            smp0 = smp[0];
        }
        for (int l = lane; l < L; l += kWarp) {
            float yl = y[l] + 1e-7f;
            float pw = 1.f;
            if (MODE == 1) pw = (y[l] > 0.f) ? table[min(l, table_len - 1)] : 0.f;
            if (MODE == 2) {
                pw = smp0 / smp[l];                                   // get_normalized_weights, dla.py:296-298
                We += yl * (sm_s0 / expf(s[l] - lse));
            }
            W += yl * pw;
        }
        W = warp_sum(W);
        if (MODE == 2) We = warp_sum(We);
        // loss + gradient
        float ll = 0.f, dsum = 0.f, lle = 0.f, dsum_e = 0.f;
        for (int l = lane; l < L; l += kWarp) {
            float yl = y[l] + 1e-7f;
            float pw = 1.f;
            if (MODE == 1) pw = (y[l] > 0.f) ? table[min(l, table_len - 1)] : 0.f;
            if (MODE == 2) pw = smp0 / smp[l];
            float w = yl * pw;
            float d = (W != 0.f) ? w / W : 0.f;                       // nan_to_num(w / W)
            float lsm = s[l] - lse;
            ll = fmaf(-d, lsm, ll);
            dsum += d;
            if (MODE == 2) {
                float we = yl * (sm_s0 / expf(lsm));
                float de = (We != 0.f) ? we / We : 0.f;
                lle = fmaf(-de, prop[l] - lse_p, lle);
                dsum_e += de;
            }
        }
        ll = warp_sum(ll);
        dsum = warp_sum(dsum);
        if (MODE == 2) {
            lle = warp_sum(lle);
            dsum_e = warp_sum(dsum_e);
        }
        for (int l = lane; l < L; l += kWarp) {
            float yl = y[l] + 1e-7f;
            float pw = 1.f;
            if (MODE == 1) pw = (y[l] > 0.f) ? table[min(l, table_len - 1)] : 0.f;
            if (MODE == 2) pw = smp0 / smp[l];
            float w = yl * pw;
            float d = (W != 0.f) ? w / W : 0.f;
            float lsm = s[l] - lse;
            dscores[(size_t)b * L + l] = (expf(lsm) * dsum - d) * W;
            if (MODE == 2) {
                float we = yl * (sm_s0 / expf(lsm));
                float de = (We != 0.f) ? we / We : 0.f;
                acc[l] += (smp[l] * dsum_e - de) * We;
            }
        }
        num += ll * W;
        den += W;
        if (MODE == 2) {
            num_e += lle * We;
            den_e += We;
        }
    }
    // block partials: [num, den, num_e, den_e, gprop[L]]
    const int width = (MODE == 2) ? 4 + L : 2;
    __shared__ float red[8][4];
    if (lane == 0) {
        red[wid][0] = num; red[wid][1] = den; red[wid][2] = num_e; red[wid][3] = den_e;
    }
    __syncthreads();
    float* mine = partials + (size_t)blockIdx.x * width;
    if (threadIdx.x < ((MODE == 2) ? 4 : 2)) {
        float s = 0.f;
        for (int q = 0; q < nw; ++q) s += red[q][threadIdx.x];
        mine[threadIdx.x] = s;
    }
    if (MODE == 2) {
        float* accs = sm + 2 * L;
        for (int l = threadIdx.x; l < L; l += blockDim.x) {
            float s = 0.f;
            for (int q = 0; q < nw; ++q) s += accs[(size_t)q * L + l];
            mine[4 + l] = s;
        }
    }
    if (last_block_ticket(counter, gridDim.x)) {
        if (MODE != 2) {
            reduce_partials(partials, gridDim.x, width, 2, sums);
        } else {
            // sums[4], then chain the exam gradient through ELU and the Linear(L,1): dprop[l] = g_l * ELU'(pre_l),
            // dprop[L] = sum_l dprop[l]   (dla.py:24-48)
            __shared__ float gsum[256];
            reduce_partials(partials, gridDim.x, width, 4, sums);
            float local = 0.f;
            const float pb = prop_b[0];
            for (int l = threadIdx.x; l < L; l += blockDim.x) {
                float g = 0.f;
                for (int b = 0; b < (int)gridDim.x; ++b) g += partials[(size_t)b * width + 4 + l];
                float pre = prop_w[l] + pb;
                float gp = g * (pre > 0.f ? 1.f : expf(pre));
                dprop[l] = gp;
                local += gp;
            }
            gsum[threadIdx.x] = local;
            __syncthreads();
            if (threadIdx.x == 0) {
                float s = 0.f;
                for (int q = 0; q < (int)blockDim.x; ++q) s += gsum[q];
                dprop[L] = s;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3: pairwise losses, one CTA per list
// ------------------------------------------------------------------------------------------------
// 1 / x as ONE rcp.approx.ftz (<= 1 ulp; the IEEE division expands to ~8 instructions).  Used on the pair path of K3,
// where every unordered pair needs four reciprocals of values in [1, 2].
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float safe_div_f(float n, float d) { return d == 0.f ? 0.f : n / d; }

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float s = 0.f;
    for (int q = 0; q < nw; ++q) s += red[q];
    return s;
}

// KIND 0: PairDebias (pairwise_debias.py:142-157)       - positions are display positions
// KIND 1: LambdaRank (lambda_rank.py:116-135, 247-291) - positions are predicted ranks (LAMBDA below)
// KIND 2: PRSrank    (prs_rank.py:94-151)              - predicted ranks; pair (r < s) weighted by delta-NDCG * ipw_r / ipw_s
//         (ipw = IPW table entry of the document's DISPLAY position; t_plus carries the table, t_minus is unused), loss =
//         weighted binary cross-entropy of p_rs = sigmoid(sigma (s_r - s_s)) against P_rs = (1 + clamp(y_r - y_s)) / 2
//
// Every UNORDERED pair {i, j} is evaluated exactly once (the reference's [L, L] matrices hold each pair twice, as
// (i, j) and (j, i), and both directions share every transcendental): thread i visits the partners j = (i + k) mod L
// for k = 1 .. L/2 ("round-robin tournament"), so within one step k the partners of a warp's lanes are all different
// and the partner-side contributions go to per-warp shared-memory arrays without atomics; a __syncwarp() per step
// is the only synchronisation.  Owner-side sums stay in registers.  Everything is added in a fixed order, so the
// result is bitwise reproducible.  Issue-bound: ~8 MUFU + ~60 FP32 instructions per pair.
// Warp w owns rows 32 * (w % nrb) .. + 31 and the k-group w / nrb: a list of 200 positions runs on 14 warps
// (7 row blocks x 2 groups of 50 steps), two lists per SM, so the long dependency chain of one pair is hidden by
// other warps.
struct PairAcc {
    float Tp, Tm, g, ls;
};

template <int KIND>
__global__ void __launch_bounds__(512, 2) pairwise_kernel(const float* __restrict__ scores,
                                                        const float* __restrict__ labels, int B, int L, float sigma,
                                                        const float* __restrict__ t_plus,
                                                        const float* __restrict__ t_minus,
                                                        float* __restrict__ dscores, float* __restrict__ out,
                                                        unsigned int* counter, float* __restrict__ partials,
                                                        int n_groups, int table_len) {
    constexpr bool LAMBDA = KIND != 0;      // ranked by predicted score
    constexpr bool PRS = KIND == 2;
    extern __shared__ float sm[];
    float* ps = sm;              // scores in position order (sorted for LAMBDA)
    float* ys = ps + L;          // labels in position order
    float* gn = ys + L;          // LAMBDA: gain 2^y - 1 ; else unused
    float* dc = gn + L;          // LAMBDA: 1/log2(pos+2)
    float* tp = dc + L;
    float* tm = tp + L;
    float* rtp = tm + L;         // 1 / t+ , 1 / t-
    float* rtm = rtp + L;
    float* accp = rtm + L;       // block accumulators of T+ / T- over this block's lists
    float* accm = accp + L;
    float* gr = accm + L;        // owner-side results of the current list: gradient, T+, T-
    float* oTp = gr + L;
    float* oTm = oTp + L;
    float* raw_s = oTm + L;      // LAMBDA: unsorted copy
    float* raw_y = raw_s + L;
    int* pos = reinterpret_cast<int*>(raw_y + L);   // LAMBDA: rank of original index
    int* ipos = pos + L;                            // LAMBDA: ideal rank (by label) of original index
    float* part = reinterpret_cast<float*>(ipos + L);  // [nwarps][3][L] partner-side g, T+, T-
    __shared__ float red[32];
    griddep_launch();
    griddep_wait();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* pg = part + (size_t)wid * 3 * L;
    float* pTp = pg + L;
    float* pTm = pTp + L;

    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        if (!PRS) {
            const float a = t_plus[i], b = t_minus[i];
            tp[i] = a;
            tm[i] = b;
            rtp[i] = 1.f / a;
            rtm[i] = 1.f / b;
        }
        accp[i] = 0.f;
        accm[i] = 0.f;
        if (LAMBDA) dc[i] = 1.f / log2f((float)i + 2.f);
    }
    const int half = L / 2;
    const int nrb = (L + 31) / 32;                         // row blocks of 32 positions
    const int kgroup = n_groups > 1 ? wid / nrb : 0;       // this warp's k-group (n_groups > 1: nw == nrb * n_groups)
    const int ksteps = (half + n_groups - 1) / n_groups;
    const int k_lo = kgroup * ksteps + 1;
    const int k_hi = min(half, (kgroup + 1) * ksteps);
    float loss_acc = 0.f, idcg_acc = 0.f;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        const float* s = scores + (size_t)b * L;
        const float* y = labels + (size_t)b * L;
        for (int q = threadIdx.x; q < 3 * L * nw; q += blockDim.x) part[q] = 0.f;
        if (LAMBDA) {
            for (int i = threadIdx.x; i < L; i += blockDim.x) {
                raw_s[i] = s[i];
                raw_y[i] = y[i];
                pos[i] = 0;
                ipos[i] = 0;
            }
            __syncthreads();
            // stable descending rank by counting (torch.sort(descending=True), lambda_rank.py:116) and the ideal
            // rank of each label for the IDCG (lambda_rank.py:126, 263-266: natural log, summed over the batch).
            // The k-groups split the j range; integer atomics are order independent.
            {
                const int jn = (L + n_groups - 1) / n_groups;
                const int j_lo = kgroup * jn, j_hi = min(L, j_lo + jn);
                for (int i = (n_groups > 1 ? (wid % nrb) * 32 + lane : (int)threadIdx.x); i < L;
                     i += (n_groups > 1 ? L : (int)blockDim.x)) {
                    const float si = raw_s[i], yi = raw_y[i];
                    int r = 0, ir = 0;
                    for (int j = j_lo; j < j_hi; ++j) {
                        const float sj = raw_s[j], yj = raw_y[j];
                        r += (sj > si) || (sj == si && j < i);
                        ir += (yj > yi) || (yj == yi && j < i);
                    }
                    if (n_groups > 1) {
                        atomicAdd(&pos[i], r);
                        atomicAdd(&ipos[i], ir);
                    } else {
                        pos[i] = r;
                        ipos[i] = ir;
                    }
                }
            }
            __syncthreads();
            for (int i = threadIdx.x; i < L; i += blockDim.x) {
                const float si = raw_s[i], yi = raw_y[i];
                const int r = pos[i];
                ps[r] = si;
                ys[r] = yi;
                const float gain = exp2f(yi) - 1.f;
                gn[r] = gain;
                if (PRS) {
                    // ipw of the document's display position i (getPropensityForOneList(use_non_clicked_data=True),
                    // propensity_estimator.py:22-42) and pw = _safe_div(1, ipw) (prs_rank.py:126)
                    const float w = t_plus[min(i, table_len - 1)];
                    tp[r] = w;
                    rtp[r] = w == 0.f ? 0.f : 1.f / w;
                }
                idcg_acc += gain / logf((float)ipos[i] + 2.f);
            }
        } else {
            for (int i = threadIdx.x; i < L; i += blockDim.x) {
                ps[i] = s[i];
                ys[i] = y[i];
            }
        }
        __syncthreads();
        // all lanes of a warp run the same number of steps (the __syncwarp below), also those without a row
        for (int ib = (n_groups > 1 ? wid % nrb : wid) * 32; ib < L; ib += (n_groups > 1 ? L + 32 : (int)blockDim.x)) {
            const int i = ib + lane;
            const bool have = i < L;
            const int ii = have ? i : 0;
            const float si = ps[ii], yi = ys[ii], tpi = tp[ii], tmi = tm[ii], rtpi = rtp[ii], rtmi = rtm[ii];
            const float gi = LAMBDA ? gn[ii] : 0.f, di = LAMBDA ? dc[ii] : 0.f;
            float Tp = 0.f, Tm = 0.f, g = 0.f, ls = 0.f;
            for (int k = k_lo; k <= k_hi; ++k) {
                int j = ii + k;
                if (j >= L) j -= L;
                // for even L the antipodal pair (k == L/2) would be met from both ends
                const bool act = have && !(2 * k == L && i >= half);
                if (PRS) {
                    const float delta = fabsf(gi - gn[j]) * fabsf(di - dc[j]);
                    if (act && delta != 0.f) {
                        // the pair counts once, as (a, b) with a ranked above b (triu(..., diagonal=1), prs_rank.py:131)
                        const bool own_first = ii < j;
                        const float sa = own_first ? si : ps[j], sb = own_first ? ps[j] : si;
                        const float ya = own_first ? yi : ys[j], yb = own_first ? ys[j] : yi;
                        const float w = delta * (own_first ? tpi * rtp[j] : tp[j] * rtpi);     // delta-NDCG * ipw_a * pw_b
                        const float d = sigma * (sa - sb);
                        const float p = rcp_fast(expf(-d) + 1.f);                             // prs_rank.py:139
                        const float t = 0.5f * (1.f + fminf(fmaxf(ya - yb, -1.f), 1.f));
                        // F.binary_cross_entropy clamps both logarithms at -100 and divides by max(p (1 - p), 1e-12)
                        // in its backward pass
                        const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);
                        ls -= w * (t * lp + (1.f - t) * l1p);
                        const float pq = p * (1.f - p);
                        const float ga = w * sigma * (p - t) * (pq / fmaxf(pq, 1e-12f));
                        g += own_first ? ga : -ga;
                        pg[j] += own_first ? -ga : ga;
                    }
                } else if (LAMBDA) {
                    const float delta = fabsf(gi - gn[j]) * fabsf(di - dc[j]);
                    if (act && delta != 0.f) {
                        const float d = sigma * (si - ps[j]);
                        // p_ij = 1 / (exp(-d) + 1), p_ji = 1 / (exp(d) + 1) from ONE exponential of -|d| (no overflow)
                        const float e = exp_ftz(-fabsf(d));
                        const float p_big = rcp_fast(1.f + e), p_small = e * p_big;
                        const float pij = d >= 0.f ? p_big : p_small, pji = d >= 0.f ? p_small : p_big;
                        const float Sij = fminf(fmaxf(yi - ys[j], -1.f), 1.f);
                        const float Pij = 0.5f * (1.f + Sij), Pji = 0.5f * (1.f - Sij);
                        // BCEWithLogits applied to the probability p (lambda_rank.py:128): p - p*P + log1p(exp(-p));
                        // its derivative needs sigmoid(p) = 1 / (1 + exp(-p)) of the same exponential
                        const float uij = exp_ftz(-pij), uji = exp_ftz(-pji);
                        const float tij = delta * (pij - pij * Pij + __logf(1.f + uij));
                        const float tji = delta * (pji - pji * Pji + __logf(1.f + uji));
                        const float sgij = rcp_fast(1.f + uij), sgji = rcp_fast(1.f + uji);
                        const float tpj = tp[j], tmj = tm[j], rtpj = rtp[j], rtmj = rtm[j];
                        const float inv_ij = (tpi * tmj == 0.f) ? 0.f : rtpi * rtmj;      // metrics._safe_div
                        const float inv_ji = (tpj * tmi == 0.f) ? 0.f : rtpj * rtmi;
                        Tp = fmaf(tij, rtmj, Tp);                 // T+_i += t_ij / t-_j
                        Tm = fmaf(tji, rtpj, Tm);                 // T-_i += t_ji / t+_j
                        ls += tij * inv_ij + tji * inv_ji;
                        const float aij = delta * (sgij - Pij) * sigma * pij * (1.f - pij) * inv_ij;
                        const float aji = delta * (sgji - Pji) * sigma * pji * (1.f - pji) * inv_ji;
                        g += aij - aji;
                        pg[j] += aji - aij;
                        pTp[j] = fmaf(tji, rtmi, pTp[j]);         // T+_j += t_ji / t-_i
                        pTm[j] = fmaf(tij, rtpi, pTm[j]);         // T-_j += t_ij / t+_i
                    }
                } else {
                    const float cj = ys[j];
                    const float mij = fminf(1.f, fmaxf(yi - cj, 0.f));
                    const float mji = fminf(1.f, fmaxf(cj - yi, 0.f));
                    if (act && (mij != 0.f || mji != 0.f)) {
                        // at most one direction is active: a = the clicked side, b = the other one
                        const bool fwd = mij != 0.f;
                        const float m = fwd ? mij : mji;
                        const float dlt = fwd ? ps[j] - si : si - ps[j];          // s_b - s_a
                        const float t = m * softplus_f(dlt);
                        const float sg = m * sigmoid_f(dlt);
                        const float rtpa = fwd ? rtpi : rtp[j], rtma = fwd ? rtmi : rtm[j];
                        const float rtpb = fwd ? rtp[j] : rtpi, rtmb = fwd ? rtm[j] : rtmi;
                        (void)rtma; (void)rtpb;
                        const float inv = rtpa * rtmb;                            // 1 / (t+_a t-_b), no safe-div
                        ls = fmaf(t, inv, ls);
                        const float ga = -sg * inv;
                        if (fwd) {
                            Tp = fmaf(t, rtmb, Tp);               // T+_a += t / t-_b
                            pTm[j] = fmaf(t, rtpa, pTm[j]);       // T-_b += t / t+_a
                            g += ga;
                            pg[j] -= ga;
                        } else {
                            pTp[j] = fmaf(t, rtmb, pTp[j]);
                            Tm = fmaf(t, rtpa, Tm);
                            pg[j] += ga;
                            g -= ga;
                        }
                    }
                }
                __syncwarp();
            }
            if (have) {
                // owner-side sums: k-group 0 writes the per-position slots, the other groups add theirs into their
                // warp's partner-side arrays (the k loop is over, every lane touches only its own index)
                if (kgroup == 0) {
                    gr[i] = g;
                    oTp[i] = Tp;
                    oTm[i] = Tm;
                } else {
                    pg[i] += g;
                    pTp[i] += Tp;
                    pTm[i] += Tm;
                }
            }
            loss_acc += ls;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < L; i += blockDim.x) {
            float g = gr[i], Tp = oTp[i], Tm = oTm[i];
            for (int w = 0; w < nw; ++w) {
                const float* q = part + (size_t)w * 3 * L;
                g += q[i];
                Tp += q[L + i];
                Tm += q[2 * L + i];
            }
            gr[i] = g;
            accp[i] += Tp;
            accm[i] += Tm;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < L; i += blockDim.x)
            dscores[(size_t)b * L + i] = LAMBDA ? gr[pos[i]] : gr[i];
    }
    __syncthreads();
    const int width = 2 * L + 2;
    float* mine = partials + (size_t)blockIdx.x * width;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        mine[i] = accp[i];
        mine[L + i] = accm[i];
    }
    const float lsum = block_sum(loss_acc, red);
    const float isum = block_sum(idcg_acc, red);
    if (threadIdx.x == 0) {
        mine[2 * L] = lsum;
        mine[2 * L + 1] = isum;
    }
    if (last_block_ticket(counter, gridDim.x)) reduce_partials(partials, gridDim.x, width, LAMBDA ? width : width - 1, out);
}

__global__ void em_update_kernel(float* __restrict__ t_plus, float* __restrict__ t_minus,
                                 const float* __restrict__ out, int L, float em_step, float expo, int safe) {
    griddep_launch();
    griddep_wait();
    const float Tp0 = out[0], Tm0 = out[L];
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        float rp = safe ? safe_div_f(out[i], Tp0) : out[i] / Tp0;
        float rm = safe ? safe_div_f(out[L + i], Tm0) : out[L + i] / Tm0;
        t_plus[i] = (1.f - em_step) * t_plus[i] + em_step * powf(rp, expo);
        t_minus[i] = (1.f - em_step) * t_minus[i] + em_step * powf(rm, expo);
    }
}

}  // namespace ub200

using namespace ub200;

extern "C" UB200_API size_t ub200_loss_workspace_bytes(int B, int L) {
    (void)B;
    return loss_ws_bytes(4 + (L > 0 ? L : 0));
}

extern "C" UB200_API size_t ub200_pair_workspace_bytes(int B, int L) {
    (void)B;
    return loss_ws_bytes(2 * (L > 0 ? L : 0) + 2);
}

static int loss_grid(int B, int lists_per_block, int per_sm = 2) {
    int g = (B + lists_per_block - 1) / lists_per_block;
    if (g > per_sm * kNumSMs) g = per_sm * kNumSMs;
    return g < 1 ? 1 : g;
}

extern "C" UB200_API int ub200_softmax_ce(const float* scores, const float* labels, int B, int L, int weight_mode,
                                const float* table, int table_len, float* dscores, float* sums, void* workspace,
                                size_t workspace_bytes, void* stream) {
    UB_CHECK(B > 0 && L > 0, 1, "softmax_ce: bad B=%d L=%d", B, L);
    UB_CHECK(scores && labels && dscores && sums && workspace, 2, "softmax_ce: null pointer");
    UB_CHECK(weight_mode == 0 || (weight_mode == 1 && table && table_len > 0), 1, "softmax_ce: bad weight_mode %d",
             weight_mode);
    UB_CHECK(workspace_bytes >= loss_ws_bytes(2), 3, "softmax_ce: workspace too small");
    LossWs w = loss_ws(workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (L <= 256) {
        // register-resident lists (the normal case: list lengths 10-200): G lanes per list, EPL elements per lane
        const int G = L <= 64 ? 8 : (L <= 128 ? 16 : 32);
        const int epl = (L + G - 1) / G;
        const int lists_per_block = 8 * (32 / G);
        int grid = (B + lists_per_block - 1) / lists_per_block;
        if (grid > kLossBlocksWide) grid = kLossBlocksWide;
#define UB_K2(MODE_, G_, EPL_)                                                                                       \
        launch_k(softmax_ce_reg_kernel<MODE_, G_, EPL_>, grid, 256, 0, st, scores, labels, B, L, table, table_len,  \
                 dscores, sums, w.counter, w.partials)
#define UB_K2_G8(MODE_)                                                                                                \
        switch (epl) {                                                                                                 \
            case 1: UB_K2(MODE_, 8, 1); break;                                                                         \
            case 2: UB_K2(MODE_, 8, 2); break;                                                                         \
            case 3: UB_K2(MODE_, 8, 3); break;                                                                         \
            case 4: UB_K2(MODE_, 8, 4); break;                                                                         \
            case 5: UB_K2(MODE_, 8, 5); break;                                                                         \
            case 6: UB_K2(MODE_, 8, 6); break;                                                                         \
            default: UB_K2(MODE_, 8, 8); break;                                                                        \
        }
        if (weight_mode == 0) {
            if (G == 8) { UB_K2_G8(0) } else if (G == 16) { if (epl <= 6) UB_K2(0, 16, 6); else UB_K2(0, 16, 8); }
            else { if (epl <= 6) UB_K2(0, 32, 6); else UB_K2(0, 32, 8); }
        } else {
            if (G == 8) { UB_K2_G8(1) } else if (G == 16) { if (epl <= 6) UB_K2(1, 16, 6); else UB_K2(1, 16, 8); }
            else { if (epl <= 6) UB_K2(1, 32, 6); else UB_K2(1, 32, 8); }
        }
#undef UB_K2_G8
#undef UB_K2
        UB_LAUNCH_CHECK("softmax_ce_reg_kernel");
        return 0;
    }
    const int grid = loss_grid(B, 8);
    if (weight_mode == 0)
        launch_k(softmax_ce_kernel<0>, grid, 256, 0, st, scores, labels, B, L, nullptr, 0, nullptr, nullptr, dscores, nullptr,
                                                   sums, w.counter, w.partials);
    else
        launch_k(softmax_ce_kernel<1>, grid, 256, 0, st, scores, labels, B, L, table, table_len, nullptr, nullptr, dscores,
                                                   nullptr, sums, w.counter, w.partials);
    UB_LAUNCH_CHECK("softmax_ce_kernel");
    return 0;
}

extern "C" UB200_API int ub200_dla_loss(const float* scores, const float* clicks, int B, int L, const float* prop_w,
                              const float* prop_b, float* dscores, float* dprop, float* sums, void* workspace,
                              size_t workspace_bytes, void* stream) {
    UB_CHECK(B > 0 && L > 0, 1, "dla_loss: bad B=%d L=%d", B, L);
    UB_CHECK(scores && clicks && prop_w && prop_b && dscores && dprop && sums && workspace, 2,
             "dla_loss: null pointer");
    UB_CHECK(workspace_bytes >= loss_ws_bytes(4 + L), 3, "dla_loss: workspace too small");
    LossWs w = loss_ws(workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (L <= 256) {
        // register-resident lists: G lanes per list, EPL positions per lane (as ub200_softmax_ce)
        // short lists: 4 lanes per list (8 lists per warp instruction, 2-step reductions, no idle register slots at L = 20)
        const int G = L <= 32 ? 4 : (L <= 64 ? 8 : (L <= 128 ? 16 : 32));
        const int epl = (L + G - 1) / G;
        const int grid = loss_grid(B, 8 * (32 / G), (G == 4 && epl <= 5) ? 3 : 2);
#define UB_DLA(G_, EPL_)                                                                                           \
        launch_k(dla_reg_kernel<G_, EPL_>, grid, 256, 0, st, scores, clicks, B, L, prop_w, prop_b, dscores, dprop, \
                 sums, w.counter, w.partials)
        if (G == 4) {
            if (epl <= 3) UB_DLA(4, 3); else if (epl <= 5) UB_DLA(4, 5); else UB_DLA(4, 8);
        } else if (G == 8) {
            if (epl <= 2) UB_DLA(8, 2); else if (epl <= 3) UB_DLA(8, 3); else if (epl <= 5) UB_DLA(8, 5); else UB_DLA(8, 8);
        } else if (G == 16) {
            if (epl <= 6) UB_DLA(16, 6); else UB_DLA(16, 8);
        } else {
            if (epl <= 6) UB_DLA(32, 6); else UB_DLA(32, 8);
        }
#undef UB_DLA
        UB_LAUNCH_CHECK("dla_reg_kernel");
        return 0;
    }
    const size_t smem = sizeof(float) * (size_t)(2 + 8) * L;
    UB_CHECK(smem <= 200 * 1024, 4, "dla_loss: list length %d too large", L);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(softmax_ce_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = loss_grid(B, 8);
    launch_k(softmax_ce_kernel<2>, grid, 256, smem, st, scores, clicks, B, L, nullptr, 0, prop_w, prop_b, dscores, dprop,
                                                  sums, w.counter, w.partials);
    UB_LAUNCH_CHECK("softmax_ce_kernel<dla>");
    return 0;
}

template <int KIND>
static int launch_pairwise(const float* scores, const float* labels, int B, int L, float sigma, const float* t_plus,
                           const float* t_minus, float* dscores, float* out, void* workspace, size_t workspace_bytes,
                           void* stream, int table_len = 0) {
    UB_CHECK(B > 0 && L > 0, 1, "pairwise: bad B=%d L=%d", B, L);
    UB_CHECK(scores && labels && t_plus && (t_minus || KIND == 2) && dscores && out && workspace, 2,
             "pairwise: null pointer");
    UB_CHECK(KIND != 2 || table_len > 0, 1, "prsrank: empty IPW table");
    UB_CHECK(workspace_bytes >= loss_ws_bytes(2 * L + 2), 3, "pairwise: workspace too small");
    LossWs w = loss_ws(workspace);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // warps = row blocks x k-groups, at most 16 (two lists per SM); lists longer than 512 loop over their row blocks
    const int nrb = (L + 31) / 32;
    int groups = 1;
    if (nrb <= 8) {
        groups = 16 / nrb;
        const int max_by_steps = (L / 2) / 4;           // at least 4 pair steps per group
        if (groups > max_by_steps) groups = max_by_steps;
        if (groups > 4) groups = 4;
        if (groups < 1) groups = 1;
    }
    int threads = 32 * nrb * groups;
    if (threads > 512) threads = 512;                   // only when groups == 1
    // 17 per-position arrays + per-warp partner-side arrays [warps][3][L]
    const size_t smem = sizeof(float) * (17 + 3 * (size_t)(threads / 32)) * (size_t)L;
    UB_CHECK(smem <= 200 * 1024, 4, "pairwise: list length %d too large", L);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(pairwise_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = loss_grid(B, 1);
    launch_k(pairwise_kernel<KIND>, grid, threads, smem, st, scores, labels, B, L, sigma, t_plus, t_minus, dscores, out,
             w.counter, w.partials, groups, table_len);
    UB_LAUNCH_CHECK("pairwise_kernel");
    return 0;
}

extern "C" UB200_API int ub200_lambdarank(const float* scores, const float* labels, int B, int L, float sigma,
                                const float* t_plus, const float* t_minus, float* dscores, float* out,
                                void* workspace, size_t workspace_bytes, void* stream) {
    return launch_pairwise<1>(scores, labels, B, L, sigma, t_plus, t_minus, dscores, out, workspace,
                              workspace_bytes, stream);
}

extern "C" UB200_API int ub200_pairdebias(const float* scores, const float* clicks, int B, int L, const float* t_plus,
                                const float* t_minus, float* dscores, float* out, void* workspace,
                                size_t workspace_bytes, void* stream) {
    return launch_pairwise<0>(scores, clicks, B, L, 1.f, t_plus, t_minus, dscores, out, workspace,
                              workspace_bytes, stream);
}

extern "C" UB200_API int ub200_prsrank(const float* scores, const float* labels, int B, int L, float sigma,
                             const float* ipw_table, int table_len, float* dscores, float* out, void* workspace,
                             size_t workspace_bytes, void* stream) {
    return launch_pairwise<2>(scores, labels, B, L, sigma, ipw_table, nullptr, dscores, out, workspace,
                              workspace_bytes, stream, table_len);
}

extern "C" UB200_API int ub200_em_update(float* t_plus, float* t_minus, const float* out, int L, float em_step, float reg_p,
                               int safe_div, void* stream) {
    UB_CHECK(L > 0 && t_plus && t_minus && out, 1, "em_update: bad arguments");
    launch_k(em_update_kernel, 1, 256, 0, static_cast<cudaStream_t>(stream), t_plus, t_minus, out, L, em_step,
                                                                      1.f / (reg_p + 1.f), safe_div);
    UB_LAUNCH_CHECK("em_update_kernel");
    return 0;
}
