// K1 - DNN ranker forward / backward for sm_100a (fp32, CUDA-core path).
//
// Replaces (reference paths): the host gather base_algorithm.py:148-152, cat+cast DNN.py:72-73, the
// [LayerNorm -> Linear -> ELU] x n + LayerNorm -> Linear(1) stack DNN.py:43-55,77 and its autograd backward.
//
// Structure per linear layer j (K_j -> N_j), rows M = L*B in position-major order:
//   forward : stats_j = rowwise (mean, rstd) of X_j ; Y_j = ELU( LN(X_j) W_j^T + c_j )   (LN fused into the A-operand load)
//   final   : one warp per row: stats + dot product with the [1,K] weight, scattered to scores[B,L]
//   backward: G_j  = dZ_j^T [xhat_j | 1]      (split over row chunks, deterministic two-stage reduction)
//             dW_j = gamma (.) G + beta (x) db, dgamma = sum_n W (.) G, dbeta = sum_n W db   (LayerNorm affine grads
//             derived from the weight-gradient GEMM, so no gradient w.r.t. the features is ever formed)
//             dXhat_j = dZ_j (W_j (.) gamma) ; dZ_{j-1} = LNbwd(dXhat_j) (.) ELU'(Y_{j-1})
#include <stdlib.h>

#include "common.cuh"
#include "mlp_f16.cuh"
#include "mlp_tc.cuh"

namespace ub200 {

// Bit mask: which GEMMs of the tensor-core friendly hidden layers run on tcgen05 (3xTF32): 1 = forward,
// 2 = data gradient, 4 = weight gradient; 0 = CUDA-core fp32 kernels everywhere.  Default 7 (env UB200_TC overrides).
// 16 = forward as fused fp16-split launches (mlp_f16.cu), 32 = fused fp16-split data-gradient chain, 64 = fp16-split
// weight gradients of all hidden layers in one launch (needs 32).
// 128 = the fp16-split forward / backward kernels leave their (already split) A-operand tiles behind as operand images and
// the weight-gradient kernel loads them with bulk copies instead of re-reading and re-converting fp32 (needs 16 | 32 | 64).
enum { TC_FWD = 1, TC_DGRAD = 2, TC_WGRAD = 4, TC_FUSED_FWD = 8, TC_F16_FWD = 16, TC_F16_BWD = 32, TC_F16_WGRAD = 64,
       TC_F16_IMG = 128, TC_ALL = 255 };
static int g_tc_mode = -1;
// activations other than ELU run through the fp32 CUDA-core kernels: the entry points pin the mask to 0 for their call
static thread_local int g_tc_override = -1;
struct TcOverride {
    int saved;
    explicit TcOverride(int m) : saved(g_tc_override) { g_tc_override = m; }
    ~TcOverride() { g_tc_override = saved; }
};
static int tc_mode() {
    if (g_tc_override >= 0) return g_tc_override;
    if (g_tc_mode < 0) {
        const char* e = getenv("UB200_TC");
        g_tc_mode = e ? atoi(e) & TC_ALL : TC_ALL;
    }
    return g_tc_mode;
}
// fp16-split path: every hidden width a multiple of 64 and at most 512, feature width a multiple of 4
static bool f16_net_ok(const LayerDims& d) {
    const int nh = d.n_layers - 1;
    if (nh < 1 || d.K[0] % 4 != 0) return false;
    for (int j = 0; j < nh; ++j)
        if (d.N[j] % 64 != 0 || d.N[j] > 512) return false;
    return true;
}
static bool f16_bwd_ok(const LayerDims& d) {
    return f16_net_ok(d) && f16::bwd_shape_ok(d.N, d.n_layers - 1);
}
// tile / row-split plan of the one-launch weight-gradient kernel (shapes only; pointers are filled in by the caller)
static void plan_wgrad16(const LayerDims& d, int M, f16::WgArgs* a) {
    memset(a, 0, sizeof(*a));
    a->n = d.n_layers - 1;
    a->M = M;
    for (int j = 0; j < a->n; ++j) {
        a->l[j].N = d.N[j];
        a->l[j].K = d.K[j];
        a->l[j].ldp = (d.K[j] + 1 + 3) / 4 * 4;
    }
    f16::wgrad_plan(a, kNumSMs);
}
// Measured (B200, profiles/r02_images_ab.txt): writing the images costs the forward / backward kernels time on their
// latency chain (the staging ring is shared with the copies), which the weight-gradient kernel only wins back when its
// conversion work is large - nets with a wide layer (c3 / c4 / c5: +6 .. +19 % per step), not DNN[256,128,64] (-7 .. -13 %).
// UB200_IMG=1 / 0 forces the images on / off for every shape the kernels support.
static bool f16_images_on(const LayerDims& d) {
    const int need = TC_F16_FWD | TC_F16_BWD | TC_F16_WGRAD | TC_F16_IMG;
    if ((tc_mode() & need) != need || !f16_bwd_ok(d)) return false;
    static const int forced = [] { const char* e = getenv("UB200_IMG"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
    if (forced >= 0) return forced == 1;
    long long widest = 0;
    for (int j = 0; j + 1 < d.n_layers; ++j) {
        const long long kn = (long long)d.K[j] * d.N[j];
        widest = kn > widest ? kn : widest;
    }
    return widest >= 100000;
}
static bool use_tc(int j, int K, int N, int what = 7) { return (tc_mode() & what) != 0 && tc_layer_ok(j, K, N); }


// ------------------------------------------------------------------------------------------------
// side streams: the weight-gradient branch of every layer (wgrad GEMM + finalize) is independent of the data-gradient
// chain once dZ_j exists, so it is forked onto one of two side streams and joined before the launcher returns.
// Under stream capture the fork/join events become edges of the CUDA graph (parallel branches).
// ------------------------------------------------------------------------------------------------
struct SideStreams {
    cudaStream_t s[2] = {nullptr, nullptr};
    cudaEvent_t fork_ev[UB200_MAX_LAYERS + 1] = {};
    cudaEvent_t done_ev[UB200_MAX_LAYERS + 1] = {};
    int device = -1;
    bool ok = false;
};
static SideStreams* side_streams() {
    static SideStreams per_dev[16];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    SideStreams* ss = &per_dev[dev];
    if (!ss->ok) {
        for (int q = 0; q < 2; ++q)
            if (cudaStreamCreateWithFlags(&ss->s[q], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        for (int q = 0; q <= UB200_MAX_LAYERS; ++q) {
            if (cudaEventCreateWithFlags(&ss->fork_ev[q], cudaEventDisableTiming) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&ss->done_ev[q], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        ss->device = dev;
        ss->ok = true;
    }
    return ss;
}
static int g_side = -1;   // env UB200_SIDE_STREAMS=0 keeps the whole backward pass on the caller's stream
static bool use_side_streams() {
    if (g_side < 0) {
        const char* e = getenv("UB200_SIDE_STREAMS");
        g_side = (e && e[0] == '0') ? 0 : 1;
    }
    return g_side == 1;
}

// ------------------------------------------------------------------------------------------------
// workspace carving
// ------------------------------------------------------------------------------------------------
struct MlpWorkspace {
    float2* stats[UB200_MAX_LAYERS];   // (mean, rstd) of the input of layer j, [M]
    float* Y[UB200_MAX_LAYERS];        // output of hidden layer j (post-ELU) [M, N_j]
    float* dz[3];                      // dZ_j rotation, [M, maxH] each (layer j reads dz[j % 3], writes dz[(j-1) % 3])
    float* dxh;                        // [M, maxH]
    float* partials[UB200_MAX_LAYERS];     // per layer: split-M partial weight gradients (layers run concurrently)
    float* fin_scratch[UB200_MAX_LAYERS];  // wgrad_finalize: per-block partial dgamma / dbeta
    unsigned int* fin_counters[UB200_MAX_LAYERS];   // wgrad_finalize: one ticket per k-tile (zero between launches)
    float* wf_hi[UB200_MAX_LAYERS];    // tensor-core path: pre-split weights (forward operand [N][Kpad])
    float* wf_lo[UB200_MAX_LAYERS];
    float* wd_hi[UB200_MAX_LAYERS];    // data-gradient operand [K][Npad] = (W * gamma)^T
    float* wd_lo[UB200_MAX_LAYERS];
    // fp16-split path (mlp_f16.cu)
    uint16_t* wf16[UB200_MAX_LAYERS];  // forward operand images of 2^8 W gamma
    uint16_t* wd16[UB200_MAX_LAYERS];  // data-gradient operand images (layers >= 1, training)
    float* bias2[UB200_MAX_LAYERS];    // b + W beta
    float* wf2;                        // final layer: gamma_F w_F [K_F] followed by c_F + beta_F . w_F [1]
    float* dz16[UB200_MAX_LAYERS];     // dZ_j [M, N_j] of every hidden layer (the fused chain produces all of them)
    unsigned int* dzmax;               // [UB200_MAX_LAYERS] running max |dZ_j| (float bits)
    // operand images for the weight-gradient kernel (mlp_f16.cuh: FwdArgs::ximg, BwdArgs::dzimg) + their bookkeeping.
    // wscale / img_bad live across steps: the workspace must start zeroed (0 = no scale yet) and stay with its shape.
    uint16_t* ximg[UB200_MAX_LAYERS];
    uint16_t* dzimg[UB200_MAX_LAYERS];
    float* wscale;                     // [UB200_MAX_LAYERS]
    unsigned int* img_bad;             // [UB200_MAX_LAYERS]
    size_t total_bytes;
};

static int round_up(int x, int a) { return (x + a - 1) / a * a; }

constexpr int kFinalBlocks = 2 * kNumSMs;   // blocks of the final-layer backward (column partial sums)

static int wgrad_splits(int M, int N, int K1) {
    int tiles = ((N + 127) / 128) * ((K1 + 127) / 128);
    int s = (2 * kNumSMs + tiles - 1) / tiles;
    int max_s = (M + 255) / 256;
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    return s;
}

static void carve(const LayerDims& d, int M, int training, char* base, MlpWorkspace* w) {
    size_t off = 0;
    int maxH = 1;
    for (int j = 0; j + 1 < d.n_layers; ++j) maxH = d.N[j] > maxH ? d.N[j] : maxH;
    for (int j = 0; j < d.n_layers; ++j) {
        w->stats[j] = reinterpret_cast<float2*>(base + off);
        off = align_up(off + sizeof(float2) * (size_t)M, 256);
    }
    if (training) {
        for (int j = 0; j + 1 < d.n_layers; ++j) {
            w->Y[j] = reinterpret_cast<float*>(base + off);
            off = align_up(off + sizeof(float) * (size_t)M * d.N[j], 256);
        }
        for (int q = 0; q < 3; ++q) {
            w->dz[q] = reinterpret_cast<float*>(base + off);
            off = align_up(off + sizeof(float) * (size_t)M * maxH, 256);
        }
        w->dxh = reinterpret_cast<float*>(base + off);
        off = align_up(off + sizeof(float) * (size_t)M * maxH, 256);
        for (int j = 0; j < d.n_layers; ++j) {
            size_t pf;
            if (j == d.n_layers - 1) {
                pf = (size_t)kFinalBlocks * (d.K[j] + 1);
            } else {
                pf = (size_t)wgrad_splits(M, d.N[j], d.K[j] + 1) * d.N[j] * (d.K[j] + 1);
                const size_t need_tc = (size_t)tc_wgrad_splits(M, d.N[j], d.K[j]) * d.N[j] * round_up(d.K[j] + 1, 4);
                pf = need_tc > pf ? need_tc : pf;
                if (f16_bwd_ok(d)) {
                    f16::WgArgs plan;
                    plan_wgrad16(d, M, &plan);
                    const size_t need16 = (size_t)plan.l[j].splits * d.N[j] * plan.l[j].ldp;
                    pf = need16 > pf ? need16 : pf;
                }
            }
            w->partials[j] = reinterpret_cast<float*>(base + off);
            off = align_up(off + sizeof(float) * pf, 256);
            w->fin_scratch[j] = reinterpret_cast<float*>(base + off);
            off = align_up(off + sizeof(float) * (size_t)((d.N[j] + 7) / 8) * 2 * d.K[j], 256);
            w->fin_counters[j] = reinterpret_cast<unsigned int*>(base + off);
            off = align_up(off + sizeof(unsigned int) * ((d.K[j] + 31) / 32), 256);
        }
    } else {
        // inference: ping-pong two activation buffers
        float* a = reinterpret_cast<float*>(base + off);
        off = align_up(off + sizeof(float) * (size_t)M * maxH, 256);
        float* b = reinterpret_cast<float*>(base + off);
        off = align_up(off + sizeof(float) * (size_t)M * maxH, 256);
        for (int j = 0; j + 1 < d.n_layers; ++j) w->Y[j] = (j & 1) ? b : a;
        w->dz[0] = w->dz[1] = w->dz[2] = w->dxh = nullptr;
        for (int j = 0; j < UB200_MAX_LAYERS; ++j) {
            w->partials[j] = w->fin_scratch[j] = nullptr;
            w->fin_counters[j] = nullptr;
        }
    }
    for (int j = 0; j < UB200_MAX_LAYERS; ++j) w->wf_hi[j] = w->wf_lo[j] = w->wd_hi[j] = w->wd_lo[j] = nullptr;
    for (int j = 0; j + 1 < d.n_layers; ++j) {
        if (!tc_layer_ok(j, d.K[j], d.N[j])) continue;
        const size_t nf = (size_t)d.N[j] * round_up(d.K[j], 32);
        w->wf_hi[j] = reinterpret_cast<float*>(base + off);
        off = align_up(off + sizeof(float) * nf, 256);
        w->wf_lo[j] = reinterpret_cast<float*>(base + off);
        off = align_up(off + sizeof(float) * nf, 256);
        if (training && j > 0) {
            const size_t nd = (size_t)d.K[j] * round_up(d.N[j], 32);
            w->wd_hi[j] = reinterpret_cast<float*>(base + off);
            off = align_up(off + sizeof(float) * nd, 256);
            w->wd_lo[j] = reinterpret_cast<float*>(base + off);
            off = align_up(off + sizeof(float) * nd, 256);
        }
    }
    for (int j = 0; j < UB200_MAX_LAYERS; ++j) {
        w->wf16[j] = w->wd16[j] = nullptr;
        w->bias2[j] = w->dz16[j] = nullptr;
    }
    w->wf2 = nullptr;
    w->dzmax = nullptr;
    w->wscale = nullptr;
    w->img_bad = nullptr;
    for (int j = 0; j < UB200_MAX_LAYERS; ++j) w->ximg[j] = w->dzimg[j] = nullptr;
    if (f16_net_ok(d)) {
        const int nh = d.n_layers - 1;
        for (int j = 0; j < nh; ++j) {
            w->wf16[j] = reinterpret_cast<uint16_t*>(base + off);
            off = align_up(off + f16::prep_bytes_wf(d.K[j], d.N[j]), 1024);
            if (training && j > 0) {
                w->wd16[j] = reinterpret_cast<uint16_t*>(base + off);
                off = align_up(off + f16::prep_bytes_wd(d.K[j], d.N[j]), 1024);
            }
            w->bias2[j] = reinterpret_cast<float*>(base + off);
            off = align_up(off + sizeof(float) * d.N[j], 256);
            if (training) {
                w->dz16[j] = reinterpret_cast<float*>(base + off);
                off = align_up(off + sizeof(float) * (size_t)M * d.N[j], 256);
            }
        }
        w->wf2 = reinterpret_cast<float*>(base + off);
        off = align_up(off + sizeof(float) * (d.K[nh] + 1), 256);
        w->dzmax = reinterpret_cast<unsigned int*>(base + off);
        off = align_up(off + sizeof(unsigned int) * UB200_MAX_LAYERS, 256);
        if (training && f16_bwd_ok(d)) {
            // (carved whenever the shapes allow it, whatever the mode mask says: the layout must not move with it)
            w->wscale = reinterpret_cast<float*>(base + off);
            off = align_up(off + sizeof(float) * UB200_MAX_LAYERS, 256);
            w->img_bad = reinterpret_cast<unsigned int*>(base + off);
            off = align_up(off + sizeof(unsigned int) * UB200_MAX_LAYERS, 256);
            for (int j = 0; j < nh; ++j) {
                off = align_up(off, 1024);
                w->ximg[j] = reinterpret_cast<uint16_t*>(base + off);
                off = align_up(off + f16::img_bytes(M, d.K[j]), 1024);
                w->dzimg[j] = reinterpret_cast<uint16_t*>(base + off);
                off = align_up(off + f16::img_bytes(M, d.N[j]), 1024);
            }
        }
    }
    w->total_bytes = off;
}

// fp16-split operand images + folded biases for this step's parameters
static int prep_f16_weights(const LayerDims& d, const MlpWorkspace& w, const float* params, int training,
                            cudaStream_t st) {
    const int nh = d.n_layers - 1;
    f16::PrepArgs t{};
    t.n = nh;
    for (int j = 0; j < nh; ++j) {
        t.W[j] = params + d.off_w[j];
        t.gamma[j] = params + d.off_g[j];
        t.beta[j] = params + d.off_b[j];
        t.bias[j] = params + d.off_c[j];
        t.wf[j] = w.wf16[j];
        t.wd[j] = training ? w.wd16[j] : nullptr;
        t.bias2[j] = w.bias2[j];
        t.K[j] = d.K[j];
        t.N[j] = d.N[j];
    }
    t.gF = params + d.off_g[nh];
    t.bF = params + d.off_b[nh];
    t.wF = params + d.off_w[nh];
    t.cF = params + d.off_c[nh];
    t.KF = d.K[nh];
    t.wf2 = w.wf2;
    t.cf2 = w.wf2 + d.K[nh];
    t.dzmax = training ? w.dzmax : nullptr;
    const bool img = training && w.wscale && f16_images_on(d);
    t.wscale = img ? w.wscale : nullptr;
    t.img_bad = img ? w.img_bad : nullptr;
    return f16::prep(t, st);
}

// split the weights of the tensor-core layers into (hi, lo) TF32 operands for this step
static int prep_tc_weights(const LayerDims& d, const MlpWorkspace& w, const float* params, int training,
                           cudaStream_t st) {
    tc::PrepTable t;
    t.n = 0;
    int max_elems = 0;
    for (int j = 0; j + 1 < d.n_layers; ++j) {
        if (!use_tc(j, d.K[j], d.N[j])) continue;
        const int q = t.n++;
        t.W[q] = params + d.off_w[j];
        t.gamma[q] = params + d.off_g[j];
        t.wf_hi[q] = w.wf_hi[j]; t.wf_lo[q] = w.wf_lo[j];
        t.wd_hi[q] = training ? w.wd_hi[j] : nullptr;
        t.wd_lo[q] = training ? w.wd_lo[j] : nullptr;
        t.K[q] = d.K[j]; t.N[q] = d.N[j];
        t.Kpad[q] = round_up(d.K[j], 32); t.Npad[q] = round_up(d.N[j], 32);
        const int e = d.N[j] * t.Kpad[q] + (t.wd_hi[q] ? d.K[j] * t.Npad[q] : 0);
        max_elems = e > max_elems ? e : max_elems;
    }
    if (t.n == 0) return 0;
    return tc_prep(t, max_elems, st);
}

// ------------------------------------------------------------------------------------------------
// row-wise kernels (one warp per row)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ const float* row_ptr(const float* X, const int32_t* docid, int r, int ld) {
    return X + (size_t)(docid ? docid[r] : r) * ld;
}

__device__ __forceinline__ float2 warp_row_stats(const float* __restrict__ x, int K, int lane) {
    float s = 0.f;
    for (int k = lane; k < K; k += kWarp) s += x[k];
    float mean = warp_sum(s) / (float)K;
    float v = 0.f;
    for (int k = lane; k < K; k += kWarp) {
        float dlt = x[k] - mean;
        v += dlt * dlt;
    }
    float var = warp_sum(v) / (float)K;
    return make_float2(mean, 1.0f / sqrtf(var + kLnEps));
}

__global__ void __launch_bounds__(256) row_stats_kernel(const float* __restrict__ X, const int32_t* __restrict__ docid,
                                                         int M, int K, float2* __restrict__ stats) {
    griddep_launch();
    griddep_wait();
    int lane = threadIdx.x & 31;
    int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= M) return;
    const float* x = row_ptr(X, docid, r, K);
    float2 st = warp_row_stats(x, K, lane);
    if (lane == 0) stats[r] = st;
}

// final layer forward: score = LN(x) . w + c, written to scores[b*L + l] for row r = l*B + b
__global__ void __launch_bounds__(256) final_fwd_kernel(const float* __restrict__ X, const int32_t* __restrict__ docid,
                                                         int M, int K, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, const float* __restrict__ w,
                                                         const float* __restrict__ c, float2* __restrict__ stats,
                                                         float* __restrict__ scores, int L, int B) {
    griddep_launch();
    griddep_wait();
    int lane = threadIdx.x & 31;
    int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= M) return;
    const float* x = row_ptr(X, docid, r, K);
    float2 st = warp_row_stats(x, K, lane);
    float acc = 0.f;
    for (int k = lane; k < K; k += kWarp) {
        float a = (x[k] - st.x) * st.y * gamma[k] + beta[k];
        acc = fmaf(a, w[k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        if (stats) stats[r] = st;
        int l = r / B, b = r - l * B;
        scores[(size_t)b * L + l] = acc + c[0];
    }
}

// final layer backward: dZ_prev[r,:] = LNbwd(ds_r * w (.) gamma) (.) ELU'(x) and per-block column partials of
// G[k] = sum_r ds_r xhat_rk, G[K] = sum_r ds_r.
__global__ void __launch_bounds__(256) final_bwd_kernel(const float* __restrict__ X, const int32_t* __restrict__ docid,
                                                         const float2* __restrict__ stats, int M, int K,
                                                         const float* __restrict__ gamma, const float* __restrict__ w,
                                                         const float* __restrict__ dscores, int L, int B,
                                                         float* __restrict__ dz_prev, float* __restrict__ colpart,
                                                         int act) {
    griddep_launch();
    griddep_wait();
    extern __shared__ float sm[];   // [8][K+1]
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* acc = sm + (size_t)wid * (K + 1);
    for (int k = lane; k <= K; k += kWarp) acc[k] = 0.f;
    __syncwarp();
    for (int r = blockIdx.x * nw + wid; r < M; r += gridDim.x * nw) {
        const float* x = row_ptr(X, docid, r, K);
        float2 st = stats[r];
        int l = r / B, b = r - l * B;
        float ds = dscores[(size_t)b * L + l];
        float s1 = 0.f, s2 = 0.f;
        for (int k = lane; k < K; k += kWarp) {
            float xh = (x[k] - st.x) * st.y;
            float dxh = ds * w[k] * gamma[k];
            s1 += dxh;
            s2 = fmaf(dxh, xh, s2);
            acc[k] = fmaf(ds, xh, acc[k]);
        }
        if (lane == 0) acc[K] += ds;
        if (dz_prev) {
            s1 = warp_sum(s1) / (float)K;
            s2 = warp_sum(s2) / (float)K;
            for (int k = lane; k < K; k += kWarp) {
                float xv = x[k];
                float xh = (xv - st.x) * st.y;
                float dxh = ds * w[k] * gamma[k];
                float dx = st.y * (dxh - s1 - xh * s2);
                dz_prev[(size_t)r * K + k] = dx * act_grad_from_out(xv, act);
            }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k <= K; k += blockDim.x) {
        float s = 0.f;
        for (int q = 0; q < nw; ++q) s += sm[(size_t)q * (K + 1) + k];
        colpart[(size_t)blockIdx.x * (K + 1) + k] = s;
    }
}

// dZ_{j-1}[r,:] = LNbwd(dXhat[r,:]) (.) ELU'(X[r,:])   (X = Y_{j-1} is both the LN input and the ELU output)
__global__ void __launch_bounds__(256) ln_bwd_elu_kernel(const float* __restrict__ dxh, const float* __restrict__ X,
                                                          const float2* __restrict__ stats, int M, int K,
                                                          float* __restrict__ dz_prev, int act) {
    griddep_launch();
    griddep_wait();
    int lane = threadIdx.x & 31;
    int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= M) return;
    const float* x = X + (size_t)r * K;
    const float* g = dxh + (size_t)r * K;
    float2 st = stats[r];
    float s1 = 0.f, s2 = 0.f;
    for (int k = lane; k < K; k += kWarp) {
        float xh = (x[k] - st.x) * st.y;
        float d = g[k];
        s1 += d;
        s2 = fmaf(d, xh, s2);
    }
    s1 = warp_sum(s1) / (float)K;
    s2 = warp_sum(s2) / (float)K;
    for (int k = lane; k < K; k += kWarp) {
        float xv = x[k];
        float xh = (xv - st.x) * st.y;
        float dx = st.y * (g[k] - s1 - xh * s2);
        dz_prev[(size_t)r * K + k] = dx * act_grad_from_out(xv, act);
    }
}

// ------------------------------------------------------------------------------------------------
// generic fp32 tiled GEMM  C[i,j] = sum_c A(i,c) * B(j,c)  with mode-specific operand loads / epilogues
// ------------------------------------------------------------------------------------------------
enum { MODE_FWD = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };

struct GemmArgs {
    int I, J, C;                 // output rows, output cols, contraction length
    const float* X;              // layer input rows [M,K] (FWD / WGRAD)
    const int32_t* docid;        // gather indices for layer 0 (or nullptr)
    const float2* stats;         // [M] (mean, rstd) of X rows
    const float* gamma;          // [K]
    const float* beta;           // [K]
    const float* W;              // [N,K]
    const float* bias;           // [N]
    const float* dZ;             // [M,N]
    float* out;
    int K, N;                    // layer dims
    int act;                     // FWD: apply ELU
    int rows_per_split;          // WGRAD
};

template <int MODE>
__device__ __forceinline__ float elem_a(const GemmArgs& a, int i, int c) {
    if (MODE == MODE_FWD) {          // i = m, c = k
        const float* x = row_ptr(a.X, a.docid, i, a.K);
        float2 st = a.stats[i];
        return (x[c] - st.x) * st.y * a.gamma[c] + a.beta[c];
    } else if (MODE == MODE_DGRAD) { // i = m, c = n
        return a.dZ[(size_t)i * a.N + c];
    } else {                         // WGRAD: i = n, c = m
        return a.dZ[(size_t)c * a.N + i];
    }
}
template <int MODE>
__device__ __forceinline__ float elem_b(const GemmArgs& a, int j, int c) {
    if (MODE == MODE_FWD) {          // j = n, c = k
        return a.W[(size_t)j * a.K + c];
    } else if (MODE == MODE_DGRAD) { // j = k, c = n
        return a.W[(size_t)c * a.K + j] * a.gamma[j];
    } else {                         // WGRAD: j = k (k == K -> ones column for the bias gradient), c = m
        if (j == a.K) return 1.f;
        const float* x = row_ptr(a.X, a.docid, c, a.K);
        float2 st = a.stats[c];
        return (x[j] - st.x) * st.y;
    }
}

template <int MODE, int BN>
__global__ void __launch_bounds__(256) gemm_kernel(GemmArgs a) {
    griddep_launch();
    griddep_wait();
    constexpr int BM = 128, BK = 16, TM = 8, TN = BN / 16;
    constexpr bool A_CC = (MODE != MODE_WGRAD);   // A contiguous in memory along the contraction index
    constexpr bool B_CC = (MODE == MODE_FWD);
    constexpr int A_PER = BM * BK / 256, B_PER = BN * BK / 256;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    int c_begin = 0, c_end = a.C;
    if (MODE == MODE_WGRAD) {
        c_begin = blockIdx.z * a.rows_per_split;
        c_end = min(a.C, c_begin + a.rows_per_split);
    }
    float acc[TM][TN];
#pragma unroll
    for (int p = 0; p < TM; ++p)
#pragma unroll
        for (int q = 0; q < TN; ++q) acc[p][q] = 0.f;
    float ra[A_PER], rb[B_PER];

    auto load_tiles = [&](int c0) {
#pragma unroll
        for (int e = 0; e < A_PER; ++e) {
            int idx = tid + e * 256;
            int i = A_CC ? idx / BK : idx % BM;
            int c = A_CC ? idx % BK : idx / BM;
            ra[e] = (i0 + i < a.I && c0 + c < c_end) ? elem_a<MODE>(a, i0 + i, c0 + c) : 0.f;
        }
#pragma unroll
        for (int e = 0; e < B_PER; ++e) {
            int idx = tid + e * 256;
            int j = B_CC ? idx / BK : idx % BN;
            int c = B_CC ? idx % BK : idx / BN;
            rb[e] = (j0 + j < a.J && c0 + c < c_end) ? elem_b<MODE>(a, j0 + j, c0 + c) : 0.f;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int e = 0; e < A_PER; ++e) {
            int idx = tid + e * 256;
            int i = A_CC ? idx / BK : idx % BM;
            int c = A_CC ? idx % BK : idx / BM;
            As[c][i] = ra[e];
        }
#pragma unroll
        for (int e = 0; e < B_PER; ++e) {
            int idx = tid + e * 256;
            int j = B_CC ? idx / BK : idx % BN;
            int c = B_CC ? idx % BK : idx / BN;
            Bs[c][j] = rb[e];
        }
    };

    if (c_begin < c_end) load_tiles(c_begin);
    for (int c0 = c_begin; c0 < c_end; c0 += BK) {
        store_tiles();
        __syncthreads();
        if (c0 + BK < c_end) load_tiles(c0 + BK);
#pragma unroll
        for (int c = 0; c < BK; ++c) {
            float av[TM], bv[TN];
#pragma unroll
            for (int p = 0; p < TM; p += 4) {
                float4 t = *reinterpret_cast<const float4*>(&As[c][ty * TM + p]);
                av[p] = t.x; av[p + 1] = t.y; av[p + 2] = t.z; av[p + 3] = t.w;
            }
#pragma unroll
            for (int q = 0; q < TN; q += 4) {
                float4 t = *reinterpret_cast<const float4*>(&Bs[c][tx * TN + q]);
                bv[q] = t.x; bv[q + 1] = t.y; bv[q + 2] = t.z; bv[q + 3] = t.w;
            }
#pragma unroll
            for (int p = 0; p < TM; ++p)
#pragma unroll
                for (int q = 0; q < TN; ++q) acc[p][q] = fmaf(av[p], bv[q], acc[p][q]);
        }
        __syncthreads();
    }

    // epilogue
    float* out = a.out;
    if (MODE == MODE_WGRAD) out += (size_t)blockIdx.z * a.I * a.J;
#pragma unroll
    for (int p = 0; p < TM; ++p) {
        int i = i0 + ty * TM + p;
        if (i >= a.I) continue;
#pragma unroll
        for (int q = 0; q < TN; ++q) {
            int j = j0 + tx * TN + q;
            if (j >= a.J) continue;
            float v = acc[p][q];
            if (MODE == MODE_FWD) {
                v += a.bias[j];
                if (a.act) v = act_fwd(v, a.act - 1);
            }
            out[(size_t)i * a.J + j] = v;
        }
    }
}

// reduce the split-M partials G[s][n][k] (k == K holds db; K1 = row stride of a partial plane) and derive all parameter
// gradients of the layer:
//   dW[n,k] = gamma_k G[n,k] + beta_k db[n];  db[n];  dgamma_k = sum_n W[n,k] G[n,k];  dbeta_k = sum_n W[n,k] db[n]
// grid (ceil(K/32), ceil(N/8)), block 32 k x 8 n.  The sums over n cross blocks: per-block partials go to `scratch`
// and the last block of each k-tile (ticket) adds them in block order, so the result is deterministic.
__global__ void __launch_bounds__(256) wgrad_finalize_kernel(const float* __restrict__ part, int S, int N, int K,
                                                              int K1, const float* __restrict__ W,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float* __restrict__ dW,
                                                              float* __restrict__ db, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, float* __restrict__ scratch,
                                                              unsigned int* __restrict__ counters) {
    griddep_launch();
    griddep_wait();
    __shared__ float sg[8][33], sb[8][33];
    __shared__ bool is_last;
    const int kx = threadIdx.x & 31, ny = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + kx, n = blockIdx.y * 8 + ny;
    const size_t plane = (size_t)N * K1;
    float g = 0.f, dbn = 0.f;
    {
        // db[n]: the 32 lanes of the warp (same n) split the S planes, then a butterfly sum (fixed order)
        const float* pn = part + (size_t)(n < N ? n : 0) * K1;
        float dpart = 0.f;
        if (n < N)
            for (int s = kx; s < S; s += 32) dpart += pn[(size_t)s * plane + K];
        dbn = warp_sum(dpart);
        if (n < N && k < K) {
            float acc[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[q] = 0.f;
            int s = 0;
            for (; s + 15 < S; s += 16) {
#pragma unroll
                for (int q = 0; q < 16; ++q) acc[q] += pn[(size_t)(s + q) * plane + k];
            }
            // ragged tail in batches of 8 / 4 / 2 / 1 with static accumulator indices (a fixed order): a scalar loop here
            // is a chain of up to 15 dependent L2 round trips (S = 27 and 40 at config 2)
            if (s + 8 <= S) {
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] += pn[(size_t)(s + q) * plane + k];
                s += 8;
            }
            if (s + 4 <= S) {
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[8 + q] += pn[(size_t)(s + q) * plane + k];
                s += 4;
            }
            if (s + 2 <= S) {
#pragma unroll
                for (int q = 0; q < 2; ++q) acc[12 + q] += pn[(size_t)(s + q) * plane + k];
                s += 2;
            }
            if (s < S) acc[14] += pn[(size_t)s * plane + k];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] += acc[q + 8];
            g = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
        }
    }
    float ag = 0.f, ab = 0.f;
    if (n < N && k < K) {
        const float wv = W[(size_t)n * K + k];
        dW[(size_t)n * K + k] = fmaf(gamma[k], g, beta[k] * dbn);
        ag = wv * g;
        ab = wv * dbn;
    }
    if (n < N && blockIdx.x == 0 && kx == 0) db[n] = dbn;
    sg[ny][kx] = ag;
    sb[ny][kx] = ab;
    __syncthreads();
    if (ny == 0 && k < K) {
        float tg = 0.f, tb = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            tg += sg[q][kx];
            tb += sb[q][kx];
        }
        scratch[((size_t)blockIdx.y * 2 + 0) * K + k] = tg;
        scratch[((size_t)blockIdx.y * 2 + 1) * K + k] = tb;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(&counters[blockIdx.x], 1u);
        is_last = (t == gridDim.y - 1);
        if (is_last) counters[blockIdx.x] = 0u;
    }
    __syncthreads();
    if (is_last) {
        // the per-block partials of this k-tile: thread row ny takes blocks ny, ny + 8, ... and the rows are combined in
        // order through shared memory - a fixed order, and an 8x shorter chain of L2 round trips than one row walking
        // all gridDim.y blocks
        __threadfence();
        float tg = 0.f, tb = 0.f;
        if (k < K)
            for (unsigned int q = ny; q < gridDim.y; q += 8) {
                tg += scratch[((size_t)q * 2 + 0) * K + k];
                tb += scratch[((size_t)q * 2 + 1) * K + k];
            }
        sg[ny][kx] = tg;
        sb[ny][kx] = tb;
        __syncthreads();
        if (ny == 0 && k < K) {
            float ag2 = 0.f, ab2 = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                ag2 += sg[q][kx];
                ab2 += sb[q][kx];
            }
            dgamma[k] = ag2;
            dbeta[k] = ab2;
        }
    }
}

// N == 1 variant for the final layer: the S per-block column partials [S][K+1] of final_bwd_kernel are reduced with
// the 8 thread rows splitting S (fixed assignment + fixed-order combine => deterministic).
__global__ void __launch_bounds__(256) final_finalize_kernel(const float* __restrict__ part, int S, int K,
                                                              const float* __restrict__ w,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float* __restrict__ dW,
                                                              float* __restrict__ db, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta) {
    griddep_launch();
    griddep_wait();
    __shared__ float sg[8][33], sd[8];
    const int kx = threadIdx.x & 31, ny = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + kx;
    const int K1 = K + 1;
    float g = 0.f, d = 0.f;
    for (int s = ny; s < S; s += 8) {
        if (k < K) g += part[(size_t)s * K1 + k];
        if (kx == 0) d += part[(size_t)s * K1 + K];
    }
    sg[ny][kx] = g;
    if (kx == 0) sd[ny] = d;
    __syncthreads();
    if (ny == 0) {
        float G = 0.f, dbn = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            G += sg[q][kx];
            dbn += sd[q];
        }
        if (k < K) {
            dW[k] = fmaf(gamma[k], G, beta[k] * dbn);
            dgamma[k] = w[k] * G;
            dbeta[k] = w[k] * dbn;
        }
        if (blockIdx.x == 0 && kx == 0) db[0] = dbn;
    }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
template <int MODE>
static void launch_gemm(const GemmArgs& a, int splits, cudaStream_t st) {
    if (a.J <= 64) {
        dim3 grid((a.J + 63) / 64, (a.I + 127) / 128, splits);
        launch_k(gemm_kernel<MODE, 64>, grid, 256, 0, st, a);
    } else {
        dim3 grid((a.J + 127) / 128, (a.I + 127) / 128, splits);
        launch_k(gemm_kernel<MODE, 128>, grid, 256, 0, st, a);
    }
}

// forward pass as fused fp16-split launches: maximal runs of layers with N <= 256 are chained in one kernel (the final
// ->1 layer rides on the last run); a wider layer runs alone, split into column tiles, and hands its activations to
// the next run through HBM/L2
static int forward_f16(const LayerDims& d, const MlpWorkspace& w, const float* feats, const int32_t* docid, int L, int B,
                       const float* params, float* scores, int training, cudaStream_t st) {
    const int nh = d.n_layers - 1, M = L * B;
    if (int rc = prep_f16_weights(d, w, params, training, st)) return rc;
    const float* X = feats;
    const int32_t* idx = docid;
    bool final_done = false;
    int j = 0;
    while (j < nh) {
        f16::FwdArgs a{};
        a.M = M; a.L = L; a.B = B;
        a.K0 = d.K[j];
        a.X = X; a.docid = idx;
        a.write_acts = training ? 1 : 0;
        a.scores = scores;
        a.wf2 = w.wf2;
        a.cf2 = w.wf2 + d.K[nh];
        int run = 1;
        if (d.N[j] > 256) {
            const int N = d.N[j];
            a.bn0 = N % 256 == 0 ? 256 : (N % 192 == 0 ? 192 : (N % 128 == 0 ? 128 : 64));
        } else {
            while (j + run < nh && run < f16::MAXF && d.N[j + run] <= 256) ++run;
            a.bn0 = d.N[j];
            a.has_final = (j + run == nh) ? 1 : 0;
        }
        a.nl = run;
        for (int q = 0; q < run; ++q) {
            a.N[q] = d.N[j + q];
            a.dual[q] = d.K[j + q] > 256 ? 1 : 0;
            a.wimg[q] = w.wf16[j + q];
            a.bias2[q] = w.bias2[j + q];
            a.Y[q] = w.Y[j + q];
            a.ximg[q] = (training && w.wscale && f16_images_on(d)) ? w.ximg[j + q] : nullptr;
            if (training || (run == 1 && !a.has_final))
                if (int rc = f16::make_tmap_f32(&a.ymap[q], w.Y[j + q], (size_t)M, (size_t)d.N[j + q])) return rc;
        }
        for (int q = 0; q <= run; ++q) a.stats[q] = w.stats[j + q];
        if (int rc = f16::fwd(a, st)) return rc;
        final_done = a.has_final != 0;
        X = w.Y[j + run - 1];
        idx = nullptr;
        j += run;
    }
    if (!final_done) {
        const int K = d.K[nh];
        launch_k(final_fwd_kernel, (M + 7) / 8, 256, 0, st, X, idx, M, K, params + d.off_g[nh], params + d.off_b[nh],
                 params + d.off_w[nh], params + d.off_c[nh], w.stats[nh], scores, L, B);
        UB_LAUNCH_CHECK("final_fwd_kernel");
    }
    return 0;
}

}  // namespace ub200

using namespace ub200;

extern "C" UB200_API int ub200_set_tc_mode(int mode) {
    const int old = tc_mode();
    g_tc_mode = mode & TC_ALL;
    return old;
}

extern "C" UB200_API size_t ub200_mlp_param_count(int F, const int* hidden, int n_hidden) {
    LayerDims d;
    if (make_dims(F, hidden, n_hidden, &d)) return 0;
    return d.n_params;
}

extern "C" UB200_API size_t ub200_mlp_workspace_bytes(int L, int B, int F, const int* hidden, int n_hidden, int training) {
    LayerDims d;
    if (make_dims(F, hidden, n_hidden, &d) || L <= 0 || B <= 0) return 0;
    MlpWorkspace w;
    carve(d, L * B, training, nullptr, &w);
    return w.total_bytes;
}

static int mlp_forward_impl(const float* feats, const int32_t* docid, int L, int B, int F, const int* hidden,
                            int n_hidden, const float* params, float* scores, void* workspace, size_t workspace_bytes,
                            int training, void* stream, int activation) {
    LayerDims d;
    UB_CHECK(make_dims(F, hidden, n_hidden, &d) == 0, 1, "mlp_forward: bad layer spec");
    UB_CHECK(activation >= UB200_ACT_ELU && activation <= UB200_ACT_SIGMOID, 1, "mlp_forward: bad activation %d", activation);
    d.act = activation;
    TcOverride fp32_only(activation != UB200_ACT_ELU ? 0 : g_tc_override);
    UB_CHECK(L > 0 && B > 0, 1, "mlp_forward: bad L=%d B=%d", L, B);
    UB_CHECK(feats && params && scores && workspace, 2, "mlp_forward: null pointer");
    const int M = L * B;
    MlpWorkspace w;
    carve(d, M, training, static_cast<char*>(workspace), &w);
    UB_CHECK(w.total_bytes <= workspace_bytes, 3, "mlp_forward: workspace too small (%zu < %zu)", workspace_bytes,
             w.total_bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int row_blocks = (M + 7) / 8;
    const float* X = feats;
    const int32_t* idx = docid;
    if ((tc_mode() & TC_F16_FWD) && f16_net_ok(d)) {
        if (training && !((tc_mode() & TC_F16_BWD) && f16_bwd_ok(d)))
            if (int rc = prep_tc_weights(d, w, params, training, st)) return rc;      // TF32 images for the old backward
        return forward_f16(d, w, feats, docid, L, B, params, scores, training, st);
    }
    if (int rc = prep_tc_weights(d, w, params, training, st)) return rc;
    if (training && (tc_mode() & TC_F16_BWD) && f16_bwd_ok(d))
        if (int rc = prep_f16_weights(d, w, params, training, st)) return rc;        // fp16 images for the new backward
    // Fused tail: from the first layer j0 whose successors fit the tensor-memory plan (N_j0 <= 256, later N <= 128),
    // ONE kernel runs the rest of the forward pass with the activations staying in tensor memory.  j0 = 0 for
    // DNN[256,128,64]; j0 = 1 for the reference default DNN[512,256,128] (layer 0 runs as a per-layer GEMM).
    const int nh = d.n_layers - 1;
    int j0 = -1;
    if ((tc_mode() & TC_FUSED_FWD) && (tc_mode() & TC_FWD)) {
        if (fused_forward_ok(F, d.N, nh)) j0 = 0;
        else if (nh >= 2 && fused_forward_ok(d.N[0], d.N + 1, nh - 1)) j0 = 1;
    }
    const int n_per_layer = (j0 >= 0) ? j0 : d.n_layers;
    for (int j = 0; j < n_per_layer; ++j) {
        const int K = d.K[j], N = d.N[j];
        const float* g = params + d.off_g[j];
        const float* bt = params + d.off_b[j];
        const float* W = params + d.off_w[j];
        const float* c = params + d.off_c[j];
        if (j == d.n_layers - 1) {
            launch_k(final_fwd_kernel, row_blocks, 256, 0, st, X, idx, M, K, g, bt, W, c, w.stats[j], scores, L, B);
            UB_LAUNCH_CHECK("final_fwd_kernel");
        } else {
            launch_k(row_stats_kernel, row_blocks, 256, 0, st, X, idx, M, K, w.stats[j]);
            UB_LAUNCH_CHECK("row_stats_kernel");
            if (use_tc(j, K, N, TC_FWD)) {
                tc::TcArgs t{};
                t.M = M; t.K = K; t.N = N; t.X = X; t.docid = idx; t.stats = w.stats[j]; t.gamma = g; t.beta = bt;
                t.Bhi = w.wf_hi[j]; t.Blo = w.wf_lo[j]; t.ldb = N; t.bias = c;
                t.out = w.Y[j]; t.ldo = N;
                if (int rc = tc_forward_layer(t, st)) return rc;
            } else {
                GemmArgs a{};
                a.I = M; a.J = N; a.C = K;
                a.X = X; a.docid = idx; a.stats = w.stats[j]; a.gamma = g; a.beta = bt; a.W = W; a.bias = c;
                a.out = w.Y[j]; a.K = K; a.N = N; a.act = 1 + d.act;   // 0 = none, 1 + UB200_ACT_* otherwise
                launch_gemm<MODE_FWD>(a, 1, st);
                UB_LAUNCH_CHECK("gemm_kernel<FWD>");
            }
            X = w.Y[j];
            idx = nullptr;
        }
    }
    if (j0 >= 0) {
        tc::FusedArgs fa{};
        fa.M = M; fa.L = L; fa.B = B; fa.n_hidden = nh - j0; fa.K0 = d.K[j0];
        fa.feats = X; fa.docid = idx;            // layer j0's input: the features (j0 = 0) or Y_{j0-1}
        for (int j = j0; j < d.n_layers; ++j) {
            const int q = j - j0;
            fa.gamma[q] = params + d.off_g[j];
            fa.beta[q] = params + d.off_b[j];
            fa.stats[q] = w.stats[j];
            if (j + 1 < d.n_layers) {
                fa.N[q] = d.N[j];
                fa.bias[q] = params + d.off_c[j];
                fa.wimg_hi[q] = w.wf_hi[j];
                fa.wimg_lo[q] = w.wf_lo[j];
                fa.Y[q] = w.Y[j];
            }
        }
        fa.w_final = params + d.off_w[d.n_layers - 1];
        fa.c_final = params + d.off_c[d.n_layers - 1];
        fa.scores = scores;
        fa.write_acts = training ? 1 : 0;
        return fused_forward(fa, st);
    }
    return 0;
}

extern "C" UB200_API int ub200_mlp_forward(const float* feats, const int32_t* docid, int L, int B, int F, const int* hidden,
                                 int n_hidden, const float* params, float* scores, void* workspace,
                                 size_t workspace_bytes, int training, void* stream) {
    return mlp_forward_impl(feats, docid, L, B, F, hidden, n_hidden, params, scores, workspace, workspace_bytes, training,
                            stream, UB200_ACT_ELU);
}
extern "C" UB200_API int ub200_mlp_forward_act(const float* feats, const int32_t* docid, int L, int B, int F,
                                     const int* hidden, int n_hidden, int activation, const float* params, float* scores,
                                     void* workspace, size_t workspace_bytes, int training, void* stream) {
    return mlp_forward_impl(feats, docid, L, B, F, hidden, n_hidden, params, scores, workspace, workspace_bytes, training,
                            stream, activation);
}

static int mlp_backward_impl(const float* feats, const int32_t* docid, int L, int B, int F, const int* hidden,
                             int n_hidden, const float* params, const float* dscores, float* grads, void* workspace,
                             size_t workspace_bytes, void* stream, int activation) {
    LayerDims d;
    UB_CHECK(make_dims(F, hidden, n_hidden, &d) == 0, 1, "mlp_backward: bad layer spec");
    UB_CHECK(activation >= UB200_ACT_ELU && activation <= UB200_ACT_SIGMOID, 1, "mlp_backward: bad activation %d",
             activation);
    d.act = activation;
    TcOverride fp32_only(activation != UB200_ACT_ELU ? 0 : g_tc_override);
    UB_CHECK(L > 0 && B > 0, 1, "mlp_backward: bad L=%d B=%d", L, B);
    UB_CHECK(feats && params && dscores && grads && workspace, 2, "mlp_backward: null pointer");
    const int M = L * B;
    MlpWorkspace w;
    carve(d, M, 1, static_cast<char*>(workspace), &w);
    UB_CHECK(w.total_bytes <= workspace_bytes, 3, "mlp_backward: workspace too small (%zu < %zu)", workspace_bytes,
             w.total_bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int row_blocks = (M + 7) / 8;
    const int nl = d.n_layers;

    SideStreams* ss = use_side_streams() ? side_streams() : nullptr;
    // SM partition: while the data-gradient chain (one 128-row tile per CTA, the critical path) runs, the concurrent
    // weight-gradient branches of layers j >= 1 are sized for the SMs it leaves free, and the chain's kernels carry
    // the greatest launch priority (measured: unpartitioned, the 80-CTA data-gradient kernels of config 2 waited for
    // SMs held by weight-gradient CTAs and ran 26 / 37 us instead of 14 / 21 us).
    const int row_tiles = (M + 127) / 128;
    const int side_budget = (ss && kNumSMs - row_tiles >= 32) ? kNumSMs - row_tiles : kNumSMs;
    static int prio_hi = 0, prio_lo = 0, prio_init = 0;
    if (!prio_init) {
        int least = 0, greatest = 0;
        if (cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess) {
            prio_hi = greatest;
            prio_lo = least;
        }
        prio_init = 1;
    }
    PriorityScope chain_priority(ss ? prio_hi : 0);
    // stream of the weight-gradient branch of layer j, forked from `st` at the current point
    auto fork = [&](int j) -> cudaStream_t {
        if (!ss) return st;
        cudaEventRecord(ss->fork_ev[j], st);
        cudaStreamWaitEvent(ss->s[j & 1], ss->fork_ev[j], 0);
        return ss->s[j & 1];
    };
    auto branch_done = [&](int j) {
        if (ss) cudaEventRecord(ss->done_ev[j], ss->s[j & 1]);
    };
    auto wait_branch = [&](int j) {
        if (ss) cudaStreamWaitEvent(st, ss->done_ev[j], 0);
    };

    const bool chain16 = (tc_mode() & TC_F16_BWD) && f16_bwd_ok(d);
    // final layer (N = 1): parameter-gradient partials (+ dZ_{nl-2} into dz[(nl-2) % 3] on the per-layer path).  With the
    // fused chain this kernel only feeds the final layer's own gradients, so it runs on a side stream next to the chain.
    {
        const int j = nl - 1, K = d.K[j];
        const float* X = (j == 0) ? feats : w.Y[j - 1];
        const int32_t* idx = (j == 0) ? docid : nullptr;
        int blocks = kFinalBlocks;
        if (blocks > row_blocks) blocks = row_blocks;
        size_t smem = sizeof(float) * 8 * (size_t)(K + 1);
        UB_CHECK(smem <= 200 * 1024, 4, "mlp_backward: final-layer width %d too large", K);
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(final_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaStream_t sf = chain16 ? fork(j) : st;
        {
            PriorityScope p(chain16 && ss ? prio_lo : launch_priority());
            launch_k(final_bwd_kernel, blocks, 256, smem, sf, X, idx, w.stats[j], M, K, params + d.off_g[j],
                     params + d.off_w[j], dscores, L, B, (j == 0 || chain16) ? nullptr : w.dz[(j - 1) % 3],
                     w.partials[j], d.act);
            UB_LAUNCH_CHECK("final_bwd_kernel");
        }
        cudaStream_t sb = chain16 ? sf : fork(j);
        PriorityScope side_priority(ss ? prio_lo : 0);
        launch_k(final_finalize_kernel, (K + 31) / 32, 256, 0, sb, w.partials[j], blocks, K, params + d.off_w[j],
                                                             params + d.off_g[j], params + d.off_b[j],
                                                             grads + d.off_w[j], grads + d.off_c[j],
                                                             grads + d.off_g[j], grads + d.off_b[j]);
        UB_LAUNCH_CHECK("final_finalize_kernel");
        branch_done(j);
    }
    if (chain16) {
        // fused data-gradient chain (mlp_f16.cu): ONE launch produces dZ_j of every hidden layer; the weight-gradient
        // branches below then only need their own dZ_j
        f16::BwdArgs b{};
        const int nh = nl - 1;
        b.M = M; b.L = L; b.B = B; b.nl = nh; b.KF = d.K[nh];
        b.dscores = dscores;
        b.wf2 = w.wf2;
        for (int q = 0; q < nh; ++q) {
            b.N[q] = d.N[q];
            b.Y[q] = w.Y[q];
            b.wd[q] = w.wd16[q];
            b.dZ[q] = w.dz16[q];
            b.dzmax[q] = w.dzmax + q;
            b.dzimg[q] = (w.wscale && f16_images_on(d)) ? w.dzimg[q] : nullptr;
            if (int rc = f16::make_tmap_f32(&b.ymap[q], w.Y[q], (size_t)M, (size_t)d.N[q])) return rc;
            if (int rc = f16::make_tmap_f32(&b.dzmap[q], w.dz16[q], (size_t)M, (size_t)d.N[q])) return rc;
        }
        for (int q = 0; q <= nh; ++q) b.stats[q] = w.stats[q];
        b.wscale = w.wscale;
        b.img_bad = w.img_bad;
        if (int rc = f16::bwd(b, st)) return rc;
    }
    const bool wgrad16 = chain16 && (tc_mode() & TC_F16_WGRAD);
    if (wgrad16) {
        // ONE launch for the weight-gradient partials of every hidden layer, then the per-layer finalize kernels side
        // by side (caller's stream + the two side streams)
        f16::WgArgs wa;
        plan_wgrad16(d, M, &wa);
        for (int j = 0; j < nl - 1; ++j) {
            f16::WgLayer& l = wa.l[j];
            l.dZ = w.dz16[j];
            l.X = (j == 0) ? feats : w.Y[j - 1];
            l.docid = (j == 0) ? docid : nullptr;
            l.stats = w.stats[j];
            l.dzmax = w.dzmax + j;
            l.out = w.partials[j];
            if (w.wscale && f16_images_on(d)) {
                l.ximg = w.ximg[j];
                l.dzimg = w.dzimg[j];
                l.wscale = w.wscale + j;
                l.img_bad = w.img_bad + j;
            }
        }
        if (int rc = f16::wgrad(wa, st)) return rc;
        for (int j = nl - 2; j >= 0; --j) {
            const int K = d.K[j], N = d.N[j];
            cudaStream_t sb = (j == 0) ? st : fork(j);
            launch_k(wgrad_finalize_kernel, dim3((K + 31) / 32, (N + 7) / 8), 256, 0, sb, w.partials[j], wa.l[j].splits, N,
                     K, wa.l[j].ldp, params + d.off_w[j], params + d.off_g[j], params + d.off_b[j], grads + d.off_w[j],
                     grads + d.off_c[j], grads + d.off_g[j], grads + d.off_b[j], w.fin_scratch[j], w.fin_counters[j]);
            UB_LAUNCH_CHECK("wgrad_finalize_kernel");
            if (j > 0) branch_done(j);
        }
        wait_branch(nl - 1);
        for (int j = 1; j < nl - 1; ++j) wait_branch(j);
        return 0;
    }
    // hidden layers, last to first; dz[j % 3] holds dZ_j [M, N_j]
    for (int j = nl - 2; j >= 0; --j) {
        const int K = d.K[j], N = d.N[j];
        const float* X = (j == 0) ? feats : w.Y[j - 1];
        const int32_t* idx = (j == 0) ? docid : nullptr;
        const float* g = params + d.off_g[j];
        const float* bt = params + d.off_b[j];
        const float* W = params + d.off_w[j];
        const float* dz = chain16 ? w.dz16[j] : w.dz[j % 3];
        // ---- weight-gradient branch (side stream): G[n, k] (k == K -> db) split over row chunks, then finalize ----
        cudaStream_t sb = fork(j);
        int S_eff, ldp;
        {
        PriorityScope side_priority((ss && j > 0) ? prio_lo : launch_priority());   // layer 0's branch ends the chain
        if (use_tc(j, K, N, TC_WGRAD)) {
            const int S = tc_wgrad_splits(M, N, K, j > 0 ? side_budget : kNumSMs);
            int rps = (M + S - 1) / S;
            rps = (rps + 31) / 32 * 32;
            S_eff = (M + rps - 1) / rps;
            ldp = round_up(K + 1, 4);
            tc::TcArgs t{};
            t.M = M; t.K = K; t.N = N; t.X = X; t.docid = idx; t.stats = w.stats[j]; t.dZ = dz;
            t.out = w.partials[j]; t.ldo = ldp; t.rows_per_split = rps; t.colsum = tc_wgrad_colsum(K) ? 1 : 0;
            if (int rc = tc_wgrad_layer(t, S_eff, sb)) return rc;
        } else {
            const int S = wgrad_splits(M, N, K + 1);
            int rps = (M + S - 1) / S;
            rps = (rps + 15) / 16 * 16;
            S_eff = (M + rps - 1) / rps;
            ldp = K + 1;
            GemmArgs a{};
            a.I = N; a.J = K + 1; a.C = M;
            a.X = X; a.docid = idx; a.stats = w.stats[j]; a.dZ = dz; a.out = w.partials[j];
            a.K = K; a.N = N; a.rows_per_split = rps;
            launch_gemm<MODE_WGRAD>(a, S_eff, sb);
            UB_LAUNCH_CHECK("gemm_kernel<WGRAD>");
        }
        launch_k(wgrad_finalize_kernel, dim3((K + 31) / 32, (N + 7) / 8), 256, 0, sb,
            w.partials[j], S_eff, N, K, ldp, W, g, bt, grads + d.off_w[j], grads + d.off_c[j], grads + d.off_g[j],
            grads + d.off_b[j], w.fin_scratch[j], w.fin_counters[j]);
        UB_LAUNCH_CHECK("wgrad_finalize_kernel");
        }
        branch_done(j);
        // ---- data-gradient chain (caller's stream) ----
        if (j > 0 && !chain16) {
            // dZ_{j-1} goes to dz[(j-1) % 3], the buffer dZ_{j+2} lived in: that layer's weight-gradient branch (forked
            // two layers ago) must have finished reading it
            const bool tc_d = use_tc(j, K, N, TC_DGRAD);
            const bool fuse = tc_d && (K == 64 || K == 128 || K == 256);   // one column tile -> LN-backward in the epilogue
            if (fuse && j + 2 <= nl - 2) wait_branch(j + 2);
            if (tc_d) {
                tc::TcArgs t{};
                t.M = M; t.K = K; t.N = N; t.dZ = dz; t.Bhi = w.wd_hi[j]; t.Blo = w.wd_lo[j];
                t.ldb = K;
                if (fuse) {
                    t.fuse_lnbwd = 1; t.X = X; t.stats = w.stats[j];
                    t.out = w.dz[(j - 1) % 3]; t.ldo = K;
                } else {
                    t.out = w.dxh; t.ldo = K;
                }
                if (int rc = tc_dgrad_layer(t, st)) return rc;
            } else {
                GemmArgs b{};
                b.I = M; b.J = K; b.C = N;
                b.dZ = dz; b.W = W; b.gamma = g; b.out = w.dxh; b.K = K; b.N = N;
                launch_gemm<MODE_DGRAD>(b, 1, st);
                UB_LAUNCH_CHECK("gemm_kernel<DGRAD>");
            }
            if (!fuse) {
                if (j + 2 <= nl - 2) wait_branch(j + 2);
                launch_k(ln_bwd_elu_kernel, row_blocks, 256, 0, st, w.dxh, X, w.stats[j], M, K, w.dz[(j - 1) % 3], d.act);
                UB_LAUNCH_CHECK("ln_bwd_elu_kernel");
            }
        }
    }
    for (int j = 0; j < nl; ++j) wait_branch(j);    // join every branch
    return 0;
}

extern "C" UB200_API int ub200_mlp_backward(const float* feats, const int32_t* docid, int L, int B, int F, const int* hidden,
                                  int n_hidden, const float* params, const float* dscores, float* grads,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    return mlp_backward_impl(feats, docid, L, B, F, hidden, n_hidden, params, dscores, grads, workspace, workspace_bytes,
                             stream, UB200_ACT_ELU);
}
extern "C" UB200_API int ub200_mlp_backward_act(const float* feats, const int32_t* docid, int L, int B, int F,
                                      const int* hidden, int n_hidden, int activation, const float* params,
                                      const float* dscores, float* grads, void* workspace, size_t workspace_bytes,
                                      void* stream) {
    return mlp_backward_impl(feats, docid, L, B, F, hidden, n_hidden, params, dscores, grads, workspace, workspace_bytes,
                             stream, activation);
}
