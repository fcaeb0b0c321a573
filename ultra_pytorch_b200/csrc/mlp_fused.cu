// K1 forward, fully fused: ONE kernel runs every hidden layer [LayerNorm -> Linear -> ELU] of the DNN ranker plus the
// final LayerNorm -> Linear(1) for a tile of 128 rows, on tcgen05 (3xTF32) with the activations staying on chip:
//
//   layer 0      A operand: features gathered from HBM, LayerNorm statistics computed in the kernel prologue,
//                normalised + (hi, lo)-split into the 128B-swizzled shared-memory ring by the 16 producer warps.
//   layer j > 0  A operand: the previous layer's activations are read back from TENSOR MEMORY (where the epilogue left
//                them), normalised with the row statistics the epilogue just reduced, split and stored to the ring -
//                they never travel through HBM/L2 on the way to the next GEMM.
//   all layers   B operand: pre-split, pre-swizzled weight images through the TMA engine (cp.async.bulk + mbarrier).
//   epilogue j   accumulators (main + correction) -> bias + ELU -> (training) Y_j to HBM for the backward pass ->
//                activations written back to tensor memory -> row mean / rstd (shifted sums per warp, Chan combine).
//   final layer  dot product of the normalised last activations with the [1, K] weight, scores scattered to [B, L].
//
// Tensor-memory plan (512 columns): layer 0 owns [0, 2 N_0); layer j >= 1 owns the half [256, 512) (j odd) or [0, 256)
// (j even).  A layer's activations live in the first N_j columns of its region; the next layer accumulates in the other
// half, so the constraints are N_0 <= 256 and N_j <= 128 for j >= 1 (DNN[256,128,64] of BASELINE config 2 fits; the
// reference default [512,256,128] uses the per-layer kernels of mlp_tc.cu).
#include "common.cuh"
#include "mlp_tc.cuh"
#include "tc_ptx.cuh"

namespace ub200 {
namespace tc {

constexpr int F_NPROD = 512;                 // warps 0-15: producers + epilogue
constexpr int F_MMA_WARP = F_NPROD / 32;     // warp 16
constexpr int F_NTHREADS = F_NPROD + 32;
constexpr int F_MAX_STAGES = 4;
constexpr int F_A_BYTES = 128 * 32 * 4;      // 16 KB per (hi | lo)
constexpr int F_RING_BYTES = 192 * 1024;     // operand ring; per layer: stages of 32 KB (A) + N * 256 B (B), 2..4 deep

__device__ __forceinline__ float4 f_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void f_st_split(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, float4 v) {
    float4 h, l;
    split_tf32(v.x, h.x, l.x);
    split_tf32(v.y, h.y, l.y);
    split_tf32(v.z, h.z, l.z);
    split_tf32(v.w, h.w, l.w);
    *reinterpret_cast<float4*>(hi_base + off) = h;
    *reinterpret_cast<float4*>(lo_base + off) = l;
}
__device__ __forceinline__ float f_elu(float z) {
    return z > 0.f ? z : __expf(z) - 1.f;
}

#ifdef UB200_TC_TIMELINE
__device__ unsigned long long g_fz_timeline[64];
#define FZ_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_fz_timeline[i] = t_; } } while (0)
#else
#define FZ_STAMP(i) do { } while (0)
#endif

__global__ void __launch_bounds__(F_NTHREADS, 1) fwd_fused_kernel(FusedArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ctl = smem + F_RING_BYTES;
    uint64_t* full_bars = reinterpret_cast<uint64_t*>(ctl);           // [layer][4] (every layer has its own ring phases)
    uint64_t* empty_bars = full_bars + UB200_MAX_LAYERS * F_MAX_STAGES;
    uint64_t* accum_bar = empty_bars + UB200_MAX_LAYERS * F_MAX_STAGES;   // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
    float* s_bias = reinterpret_cast<float*>(ctl + 640);              // [256] bias of the current layer
    float* s_gamma = s_bias + 256;                                    // [256] LayerNorm weight of the current layer's input
    float* s_beta = s_gamma + 256;                                    // [256] LayerNorm bias   of the current layer's input
    float* s_part = s_beta + 256;                                     // [4][128][3] per-warp-group row partials

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    griddep_launch();
    FZ_STAMP(0);
    const int i0 = blockIdx.x * 128;
    const int q4 = warp & 3, cg = (warp >> 2) & 3;
    const int trow = q4 * 32 + lane;                                  // this thread's row inside the tile (TMEM lane)
    const int grow = i0 + trow;

    if (tid == 0) {
        for (int s = 0; s < UB200_MAX_LAYERS * F_MAX_STAGES; ++s) {
            mbar_init(&full_bars[s], F_NPROD / 32 + 1);     // one arrival per producer warp + the TMA issuer
            mbar_init(&empty_bars[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
    }
    if (warp == F_MMA_WARP) tmem_alloc(tmem_slot, 512);
    griddep_wait();      // barrier init + tensor-memory allocation overlap the previous kernel's tail
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tlane = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    FZ_STAMP(1);
    float row_mean = 0.f, row_rstd = 1.f;       // statistics of this thread's row for the layer being produced (j > 0)

    for (int j = 0; j < a.n_hidden; ++j) {
        const int K = (j == 0) ? a.K0 : a.N[j - 1];
        const int N = a.N[j];
        const int n_chunks = (K + 31) / 32;
        const uint32_t acc_col = (j == 0) ? 0u : ((j & 1) ? 256u : 0u);
        const uint32_t prev_col = (j <= 1) ? 0u : (((j - 1) & 1) ? 256u : 0u);   // where y_{j-1} lives
        const int b_bytes = N * 128;                                              // one (hi | lo) weight tile
        const int stage_bytes = 2 * F_A_BYTES + 2 * b_bytes;
        const int n_stages = min(F_MAX_STAGES, F_RING_BYTES / stage_bytes);
        uint64_t* full_bar = full_bars + j * F_MAX_STAGES;
        uint64_t* empty_bar = empty_bars + j * F_MAX_STAGES;

        if (warp < F_MMA_WARP) {
            // ---- stage this layer's bias [N] and (j > 0) the LayerNorm affine parameters of its input [K] ----
            for (int n = tid; n < N; n += F_NPROD) s_bias[n] = a.bias[j][n];
            if (j > 0)
                for (int k = tid; k < K; k += F_NPROD) {
                    s_gamma[k] = a.gamma[j][k];
                    s_beta[k] = a.beta[j][k];
                }
            asm volatile("bar.sync 1, %0;" ::"n"(F_NPROD) : "memory");
            // =========================== producers ===========================
            if (j == 0) {
                const int c = tid & 7;
                const float* xrow[2];
                float2 st[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int m = i0 + (tid >> 3) + e * 64;
                    xrow[e] = (m < a.M) ? a.feats + (size_t)(a.docid ? a.docid[m] : m) * K : nullptr;
                    // LayerNorm statistics of the gathered feature row in ONE pass: sums shifted by the row's first
                    // element (no E[x^2] - mean^2 cancellation); the 8 threads of a row (consecutive lanes) cover
                    // their 4-float chunks, then a 3-step butterfly inside the 8-lane group
                    float s1 = 0.f, s2 = 0.f;
                    const float shift0 = xrow[e] ? xrow[e][0] : 0.f;
                    if (xrow[e])
                        for (int cc = c * 4; cc < K; cc += 32) {
                            const float4 v = f_ld4(xrow[e] + cc);
                            const float dx = v.x - shift0, dy = v.y - shift0, dz = v.z - shift0, dw = v.w - shift0;
                            s1 += (dx + dy) + (dz + dw);
                            s2 += (dx * dx + dy * dy) + (dz * dz + dw * dw);
                        }
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 4);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, 4);
                    const float mean = shift0 + s1 / (float)K;
                    s2 = fmaxf(s2 - s1 * s1 / (float)K, 0.f);
                    st[e] = make_float2(mean, 1.0f / sqrtf(s2 / (float)K + kLnEps));
                    if (xrow[e] && c == 0 && a.write_acts) a.stats[0][m] = st[e];
                }
                float4 cur[2], nxt[2], g_cur = zero4, b_cur = zero4, g_nxt = zero4, b_nxt = zero4;
                auto load_chunk = [&](int it, float4* xv, float4& g, float4& b) {
                    const int cc = it * 32 + c * 4;
                    const bool kv = cc < K;
#pragma unroll
                    for (int e = 0; e < 2; ++e) xv[e] = (kv && xrow[e]) ? f_ld4(xrow[e] + cc) : zero4;
                    g = kv ? f_ld4(a.gamma[0] + cc) : zero4;
                    b = kv ? f_ld4(a.beta[0] + cc) : zero4;
                };
                FZ_STAMP(2);
                load_chunk(0, cur, g_cur, b_cur);
                for (int it = 0; it < n_chunks; ++it) {
                    const int s = it % n_stages;
                    const uint32_t ph = (uint32_t)(it / n_stages) & 1u;
                    if (it + 1 < n_chunks) load_chunk(it + 1, nxt, g_nxt, b_nxt);
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    uint8_t* a_hi = smem + s * stage_bytes;
                    uint8_t* a_lo = a_hi + F_A_BYTES;
                    uint8_t* b_hi = a_lo + F_A_BYTES;
                    uint8_t* b_lo = b_hi + b_bytes;
                    if (tid == 0) {
                        const uint32_t bytes = (uint32_t)N * 128u;
                        const size_t goff = (size_t)it * N * 32;
                        mbar_arrive_expect_tx(&full_bar[s], 2 * bytes);
                        bulk_g2s(b_hi, a.wimg_hi[j] + goff, bytes, &full_bar[s]);
                        bulk_g2s(b_lo, a.wimg_lo[j] + goff, bytes, &full_bar[s]);
                    }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float4 v = cur[e];
                        v.x = (v.x - st[e].x) * st[e].y * g_cur.x + b_cur.x;
                        v.y = (v.y - st[e].x) * st[e].y * g_cur.y + b_cur.y;
                        v.z = (v.z - st[e].x) * st[e].y * g_cur.z + b_cur.z;
                        v.w = (v.w - st[e].x) * st[e].y * g_cur.w + b_cur.w;
                        if (!xrow[e]) v = zero4;
                        f_st_split(a_hi, a_lo, swz128((tid >> 3) + e * 64, c), v);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[s]);
                    cur[0] = nxt[0]; cur[1] = nxt[1];
                    g_cur = g_nxt; b_cur = b_nxt;
                }
            } else {
                // A operand from tensor memory: for chunk `it` (32 activations of the previous layer) every warp reads an
                // 8-column slice of its rows back from TMEM (written by the epilogue warps of the same lane quadrant;
                // ordered by tcgen05.wait::st + tcgen05 fences around the statistics barrier)
                for (int it = 0; it < n_chunks; ++it) {
                    const int s = it % n_stages;
                    const uint32_t ph = (uint32_t)(it / n_stages) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    uint8_t* a_hi = smem + s * stage_bytes;
                    uint8_t* a_lo = a_hi + F_A_BYTES;
                    uint8_t* b_hi = a_lo + F_A_BYTES;
                    uint8_t* b_lo = b_hi + b_bytes;
                    if (tid == 0) {
                        const uint32_t bytes = (uint32_t)N * 128u;
                        const size_t goff = (size_t)it * N * 32;
                        mbar_arrive_expect_tx(&full_bar[s], 2 * bytes);
                        bulk_g2s(b_hi, a.wimg_hi[j] + goff, bytes, &full_bar[s]);
                        bulk_g2s(b_lo, a.wimg_lo[j] + goff, bytes, &full_bar[s]);
                    }
                    if (j == 1 && it == 3) FZ_STAMP(48);
                    {
                        // every warp converts an 8-column slice of its 32 rows: columns it*32 + cg*8 .. +8
                        float y[8];
                        tmem_ld8(tlane + prev_col + it * 32 + cg * 8, y);
                        if (j == 1 && it == 3) FZ_STAMP(49);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int col = it * 32 + cg * 8 + h * 4;
                            const float4 g = *reinterpret_cast<const float4*>(s_gamma + col);
                            const float4 b = *reinterpret_cast<const float4*>(s_beta + col);
                            float4 v;
                            v.x = (y[h * 4 + 0] - row_mean) * row_rstd * g.x + b.x;
                            v.y = (y[h * 4 + 1] - row_mean) * row_rstd * g.y + b.y;
                            v.z = (y[h * 4 + 2] - row_mean) * row_rstd * g.z + b.z;
                            v.w = (y[h * 4 + 3] - row_mean) * row_rstd * g.w + b.w;
                            if (grow >= a.M) v = zero4;
                            f_st_split(a_hi, a_lo, swz128(trow, cg * 2 + h), v);
                        }
                        if (j == 1 && it == 3) FZ_STAMP(50);
                        fence_proxy_async();
                        if (j == 1 && it == 3) FZ_STAMP(51);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[s]);
                    if (j == 1 && it == 3) FZ_STAMP(52);
                    if (j == 1 && it == 2) FZ_STAMP(47);
                }
            }
        } else if (lane == 0) {
            // =========================== MMA issuer ===========================
            const uint32_t idesc = make_idesc_tf32(N, 0, 0);
            for (int it = 0; it < n_chunks; ++it) {
                const int s = it % n_stages;
                const uint32_t ph = (uint32_t)(it / n_stages) & 1u;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + s * stage_bytes);
                const uint32_t a_lo = a_hi + F_A_BYTES;
                const uint32_t b_hi = a_lo + F_A_BYTES;
                const uint32_t b_lo = b_hi + b_bytes;
                const int rem = K - it * 32;
                const int nk8 = rem >= 32 ? 4 : (rem + 7) / 8;
                for (int k8 = 0; k8 < nk8; ++k8) {
                    const uint64_t da_hi = make_smem_desc(a_hi + k8 * 32, 16, 1024);
                    const uint64_t da_lo = make_smem_desc(a_lo + k8 * 32, 16, 1024);
                    const uint64_t db_hi = make_smem_desc(b_hi + k8 * 32, 16, 1024);
                    const uint64_t db_lo = make_smem_desc(b_lo + k8 * 32, 16, 1024);
                    const uint32_t acc = (it | k8) != 0 ? 1u : 0u;
                    mma_tf32(tmem_base + acc_col + N, da_lo, db_hi, idesc, acc);
                    mma_tf32(tmem_base + acc_col + N, da_hi, db_lo, idesc, 1u);
                    mma_tf32(tmem_base + acc_col, da_hi, db_hi, idesc, acc);
                }
                mma_commit(&empty_bar[s]);
            }
            mma_commit(accum_bar);
        }
        FZ_STAMP(8 + 4 * j);      // producers of layer j done
        __syncwarp();
        if (warp < F_MMA_WARP) {
            // =========================== epilogue of layer j ===========================
            mbar_wait(accum_bar, (uint32_t)j & 1u);
            __syncwarp();
            tc_fence_after();
            FZ_STAMP(9 + 4 * j);  // accumulator ready
            // pass 1: y = ELU(main + corr + bias) -> HBM (training) and back into tensor memory; shifted row sums
            float shift = 0.f, s1 = 0.f, s2 = 0.f;
            int cnt = 0;
            for (int cb = cg; cb < N / 32; cb += 4) {
                float v[32], corr[32];
                if (j == 0 && cnt == 0) FZ_STAMP(40);
                tmem_ld32x2(tlane + acc_col + cb * 32, tlane + acc_col + N + cb * 32, v, corr);
                if (j == 0 && cnt == 0) FZ_STAMP(41);
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] = f_elu(v[q] + corr[q] + s_bias[cb * 32 + q]);
                if (j == 0 && cnt == 0) FZ_STAMP(42);
                if (cnt == 0) shift = v[0];
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const float dlt = v[q] - shift;
                    s1 += dlt;
                    s2 = fmaf(dlt, dlt, s2);
                }
                if (j == 0 && cnt == 0) FZ_STAMP(43);
                tmem_st32(tlane + acc_col + cb * 32, v);
                if (j == 0 && cnt == 0) FZ_STAMP(44);
                cnt += 32;
                if (a.write_acts && grow < a.M) {
                    float* dst = a.Y[j] + (size_t)grow * N + cb * 32;
#pragma unroll
                    for (int q = 0; q < 32; q += 4)
                        *reinterpret_cast<float4*>(dst + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
                }
                if (j == 0 && cnt == 32) FZ_STAMP(45);
            }
            FZ_STAMP(10 + 4 * j);  // pass 1 done
            tmem_wait_st();
            // per-warp-group partial (count, mean, M2) of this row, combined across the 4 groups in fixed order
            {
                float pm = 0.f, pM2 = 0.f;
                if (cnt > 0) {
                    pm = shift + s1 / (float)cnt;
                    pM2 = s2 - s1 * s1 / (float)cnt;
                }
                float* p = s_part + ((size_t)cg * 128 + trow) * 3;
                p[0] = (float)cnt; p[1] = pm; p[2] = pM2;
            }
            tc_fence_before();
            asm volatile("bar.sync 1, %0;" ::"n"(F_NPROD) : "memory");
            tc_fence_after();
            {
                float n = 0.f, mean = 0.f, M2 = 0.f;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float* p = s_part + ((size_t)g * 128 + trow) * 3;
                    const float ng = p[0];
                    if (ng > 0.f) {
                        const float dlt = p[1] - mean, tot = n + ng;
                        mean += dlt * ng / tot;
                        M2 += p[2] + dlt * dlt * n * ng / tot;
                        n = tot;
                    }
                }
                row_mean = mean;
                row_rstd = 1.0f / sqrtf(fmaxf(M2, 0.f) / (float)N + kLnEps);
                if (cg == 0 && a.write_acts && grow < a.M) a.stats[j + 1][grow] = make_float2(row_mean, row_rstd);
            }
            // everybody has read s_part / s_bias before the next layer (or the final dot) overwrites them
            asm volatile("bar.sync 1, %0;" ::"n"(F_NPROD) : "memory");
            FZ_STAMP(11 + 4 * j);  // statistics done
        }
    }

    // =========================== final layer: score = LN(y_last) . w + c ===========================
    if (warp < F_MMA_WARP) {
        const int jl = a.n_hidden - 1;
        const int N = a.N[jl];
        const uint32_t y_col = (jl == 0) ? 0u : ((jl & 1) ? 256u : 0u);
        for (int n = tid; n < N; n += F_NPROD) {
            s_gamma[n] = a.gamma[a.n_hidden][n];
            s_beta[n] = a.beta[a.n_hidden][n];
            s_bias[n] = a.w_final[n];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(F_NPROD) : "memory");
        float part = 0.f;
        for (int cb = cg; cb < N / 32; cb += 4) {
            float y[32];
            tmem_ld32(tlane + y_col + cb * 32, y);
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int n = cb * 32 + q;
                const float an = (y[q] - row_mean) * row_rstd * s_gamma[n] + s_beta[n];
                part = fmaf(an, s_bias[n], part);
            }
        }
        s_part[(size_t)cg * 128 + trow] = part;
        asm volatile("bar.sync 1, %0;" ::"n"(F_NPROD) : "memory");
        if (cg == 0 && grow < a.M) {
            const float sc = ((s_part[trow] + s_part[128 + trow]) + (s_part[256 + trow] + s_part[384 + trow])) +
                             a.c_final[0];
            const int l = grow / a.B, b = grow - l * a.B;
            a.scores[(size_t)b * a.L + l] = sc;
        }
    }
    FZ_STAMP(5);
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == F_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    FZ_STAMP(6);
}

}  // namespace tc

#ifdef UB200_TC_TIMELINE
extern "C" UB200_API int ub200_fused_timeline(unsigned long long* out64) {
    return (int)cudaMemcpyFromSymbol(out64, tc::g_fz_timeline, sizeof(unsigned long long) * 64);
}
#endif

bool fused_forward_ok(int F, const int* N, int n_hidden) {
    if (n_hidden < 1 || n_hidden >= UB200_MAX_LAYERS || F % 4 != 0) return false;
    for (int j = 0; j < n_hidden; ++j) {
        if (N[j] % 64 != 0) return false;
        if (N[j] > (j == 0 ? 256 : 128)) return false;
    }
    return true;
}

int fused_forward(const tc::FusedArgs& a, cudaStream_t st) {
    constexpr int smem = tc::F_RING_BYTES + 1024 + 640 + 3 * 256 * 4 + 4 * 128 * 3 * 4 + 64;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc::fwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        UB_CHECK(e == cudaSuccess, 100, "fwd_fused_kernel attribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    launch_k(tc::fwd_fused_kernel, (a.M + 127) / 128, tc::F_NTHREADS, smem, st, a);
    UB_LAUNCH_CHECK("fwd_fused_kernel");
    return 0;
}

}  // namespace ub200
