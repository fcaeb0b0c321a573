// K1 on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate through 3xTF32 error compensation.
//
// Every dense contraction of the DNN ranker (forward  Y = LN(X) W^T,  data gradient  dXhat = dZ (W.gamma),
// weight gradient  G = dZ^T [xhat | 1]) runs through ONE warp-specialised kernel template:
//
//   warps 0-7 : operand producers. They read the fp32 activations / weights from global memory (L2), apply the
//               fused prologue (LayerNorm of the A operand for the forward pass, normalisation for the weight
//               gradient), split every value into a TF32-exact high part and an fp32 remainder
//               (x = hi + lo) and store both into 128B-swizzled shared-memory tiles of a multi-stage ring.
//   warp 8    : one elected thread issues  hi*hi + lo*hi + hi*lo  as three tcgen05.mma.kind::tf32 per K=8 step,
//               accumulating in fp32 in tensor memory; tcgen05.commit releases the ring slots / signals the epilogue.
//   warps 0-3 : epilogue: tcgen05.ld of the accumulator tile (thread = row), bias + ELU (forward) and the store.
//
// The dropped lo*lo term is O(2^-22) relative, the same order as fp32 rounding, so the results stay
// inside the 1e-5 parity bound where plain TF32 (2^-11) does not.
#include "common.cuh"
#include "mlp_tc.cuh"
#include "tc_ptx.cuh"

namespace ub200 {
namespace tc {

enum { KIND_FWD = 0, KIND_DGRAD = 1, KIND_WGRAD = 2 };

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;                 // fp32 elements per 128-byte swizzle row
constexpr int NPROD = 512;                  // producer threads (warps 0-15); all of them also run the epilogue
constexpr int MMA_WARP = NPROD / 32;        // warp 16 issues the MMAs and owns the TMEM allocation
constexpr int NTHREADS = NPROD + 32;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 4;   // 16 KB

__host__ __device__ constexpr int stage_bytes(int block_n) { return 2 * A_TILE_BYTES + 2 * block_n * BLOCK_K * 4; }
__host__ __device__ constexpr int num_stages(int block_n) {
    return (220 * 1024) / stage_bytes(block_n) > 4 ? 4 : (220 * 1024) / stage_bytes(block_n);
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st_split(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, float4 v) {
    float4 h, l;
    split_tf32(v.x, h.x, l.x);
    split_tf32(v.y, h.y, l.y);
    split_tf32(v.z, h.z, l.z);
    split_tf32(v.w, h.w, l.w);
    *reinterpret_cast<float4*>(hi_base + off) = h;
    *reinterpret_cast<float4*>(lo_base + off) = l;
}
// ELU for the forward epilogue.  Negative side: ex2.approx based exp(z) - 1; its ABSOLUTE error stays below 3e-7
// (< 1e-6 of the activation scale, the quantity the parity bound is stated in) and the backward pass uses
// ELU' = y + 1 of the value actually produced, so forward and backward stay consistent.
__device__ __forceinline__ float elu_fast(float z) {
    return z > 0.f ? z : __expf(z) - 1.f;
}

#ifdef UB200_TC_TIMELINE
__device__ unsigned long long g_tc_timeline[64];
#define TC_STAMP(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_tc_timeline[i] = t_; } } while (0)
#else
#define TC_STAMP(i) do { } while (0)
#endif

template <int KIND, int BLOCK_N>
__global__ void __launch_bounds__(NTHREADS, 1) tc_gemm_kernel(TcArgs a) {
    constexpr int STAGES = num_stages(BLOCK_N);
    constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 4;
    constexpr int STAGE_BYTES = stage_bytes(BLOCK_N);
    constexpr int NB = BLOCK_N / 32;                         // 32-column blocks of the B operand (MN-major case)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
    float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 128);   // [BLOCK_N] bias slice (FWD)

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    griddep_launch();
    if (tid == 0) TC_STAMP(0);
    const int i0 = blockIdx.x * BLOCK_M;        // first row of the MMA "M" dimension (m, or n for WGRAD)
    // WGRAD: blockIdx.y = row split, blockIdx.z = column tile
    const int j0 = (KIND == KIND_WGRAD ? blockIdx.z : blockIdx.y) * BLOCK_N;   // first row of the MMA "N" dimension
    int c_begin = 0, c_end;
    if (KIND == KIND_FWD) c_end = a.K;
    else if (KIND == KIND_DGRAD) c_end = a.N;
    else {
        c_begin = blockIdx.y * a.rows_per_split;
        c_end = min(a.M, c_begin + a.rows_per_split);
    }
    const int n_chunks = (c_end - c_begin + BLOCK_K - 1) / BLOCK_K;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], KIND == KIND_WGRAD ? NPROD / 32 : NPROD / 32 + 1);   // one arrival per warp
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
    }
    // two fp32 accumulators in tensor memory: [0, BLOCK_N) main = hi*hi, [BLOCK_N, 2*BLOCK_N) correction = lo*hi + hi*lo.
    // The tensor core truncates when it accumulates, a one-sided error proportional to the accumulator magnitude and
    // the number of accumulation steps; keeping the 2^-12-times-smaller correction terms out of the main accumulator
    // cuts its accumulation count by 3 (measured at K = 700: scores 1.1e-5 -> inside the 1e-5 bound).
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, 2 * BLOCK_N);
    griddep_wait();      // everything above is on-chip set-up; global memory is touched only from here on
    if (KIND == KIND_FWD && tid < BLOCK_N) sbias[tid] = (j0 + tid < a.N) ? a.bias[j0 + tid] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) TC_STAMP(1);

    float4 csum_keep = zero4;
    if (warp < MMA_WARP) {
        // =========================== producers ===========================
        // Global loads of chunk it+1 are issued before chunk it is transformed and stored (register double
        // buffering), so the L2 / HBM latency overlaps the MMAs of earlier chunks.
        if (KIND != KIND_WGRAD) {
            // A, K-major: 128 rows (m) x 8 chunks of 4 contraction elements; thread -> chunk c = tid & 7 of the
            // rows (tid >> 3) + 64 e.  Row pointers / LayerNorm statistics are loop invariant.
            // B (pre-split, pre-swizzled weight images) arrives through the TMA engine (cp.async.bulk).
            constexpr int RPT = BLOCK_M * 8 / NPROD;    // rows per thread = 2
            const int c = tid & 7;
            const float* xrow[RPT];
            float2 st[RPT];
#pragma unroll
            for (int e = 0; e < RPT; ++e) {
                const int m = i0 + (tid >> 3) + e * (NPROD / 8);
                xrow[e] = nullptr;
                st[e] = make_float2(0.f, 1.f);
                if (m < a.M) {
                    if (KIND == KIND_FWD) {
                        xrow[e] = a.X + (size_t)(a.docid ? a.docid[m] : m) * a.K;
                        st[e] = a.stats[m];
                    } else {
                        xrow[e] = a.dZ + (size_t)m * a.N;
                    }
                }
            }
            float4 cur[RPT], nxt[RPT], g_cur = zero4, b_cur = zero4, g_nxt = zero4, b_nxt = zero4;
            auto load_chunk = [&](int it, float4* xv, float4& g, float4& b) {
                const int cc = c_begin + it * BLOCK_K + c * 4;
                const bool kv = cc < c_end;
#pragma unroll
                for (int e = 0; e < RPT; ++e) xv[e] = (kv && xrow[e]) ? ld4(xrow[e] + cc) : zero4;
                if (KIND == KIND_FWD) {
                    g = kv ? ld4(a.gamma + cc) : zero4;
                    b = kv ? ld4(a.beta + cc) : zero4;
                }
            };
            if (tid == 0) TC_STAMP(2);
            if (n_chunks > 0) load_chunk(0, cur, g_cur, b_cur);
            for (int it = 0; it < n_chunks; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                if (it + 1 < n_chunks) load_chunk(it + 1, nxt, g_nxt, b_nxt);
                mbar_wait(&empty_bar[s], ph ^ 1u);
                if (tid == 0 && it < 8) TC_STAMP(8 + 2 * it);
                uint8_t* a_hi = smem + s * STAGE_BYTES;
                uint8_t* a_lo = a_hi + A_TILE_BYTES;
                uint8_t* b_hi = a_lo + A_TILE_BYTES;
                uint8_t* b_lo = b_hi + B_TILE_BYTES;
                if (tid == 0) {
                    // weight image: [chunk][rows x 128 B, swizzled]; this tile = rows j0 .. j0+BLOCK_N of chunk `it`
                    const size_t goff = ((size_t)it * a.ldb + j0) * BLOCK_K;
                    mbar_arrive_expect_tx(&full_bar[s], 2 * B_TILE_BYTES);
                    bulk_g2s(b_hi, a.Bhi + goff, B_TILE_BYTES, &full_bar[s]);
                    bulk_g2s(b_lo, a.Blo + goff, B_TILE_BYTES, &full_bar[s]);
                }
#pragma unroll
                for (int e = 0; e < RPT; ++e) {
                    float4 v = cur[e];
                    if (KIND == KIND_FWD) {
                        v.x = (v.x - st[e].x) * st[e].y * g_cur.x + b_cur.x;
                        v.y = (v.y - st[e].x) * st[e].y * g_cur.y + b_cur.y;
                        v.z = (v.z - st[e].x) * st[e].y * g_cur.z + b_cur.z;
                        v.w = (v.w - st[e].x) * st[e].y * g_cur.w + b_cur.w;
                        if (!xrow[e]) v = zero4;
                    }
                    st_split(a_hi, a_lo, swz128((tid >> 3) + e * (NPROD / 8), c), v);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[s]);
                if (tid == 0 && it < 8) TC_STAMP(9 + 2 * it);
#pragma unroll
                for (int e = 0; e < RPT; ++e) cur[e] = nxt[e];
                g_cur = g_nxt;
                b_cur = b_nxt;
            }
        } else {
            // MN-major operands straight from the row-major activations (coalesced 16-byte loads along n / k):
            //   A = dZ^T     : 32 contraction rows (m) x 4 blocks of 32 n x 8 chunks   -> 2 float4 per thread
            //   B = [xhat|1]^T: 32 contraction rows (m) x NB blocks of 32 k x 8 chunks  -> NB/2 float4 per thread
            // Every thread keeps a fixed 4-column group; only the contraction row advances from chunk to chunk.
            constexpr int NA = 32 * 4 * 8 / NPROD;         // 2
            constexpr int NBV = 32 * NB * 8 / NPROD;       // NB / 2
            constexpr int ROWS_B = NPROD / (8 * NB);       // contraction rows covered per e step
            const int c = tid & 7;
            const int blk_a = (tid >> 3) & 3, ml_a = tid >> 5;           // + e * 16
            const int blk_b = (tid >> 3) % NB, ml_b = tid / (8 * NB);    // + e * ROWS_B
            const int n_col = i0 + blk_a * 32 + c * 4;
            const int k_col = j0 + blk_b * 32 + c * 4;
            const bool a_ok = n_col < a.N;
            // 0 full, 1 row tail (+ ones column), 2 zero.  colsum mode (K % 64 == 0): no ones column, db comes from csum
            const int b_kind = (k_col + 3 < a.K) ? 0 : ((k_col <= a.K && !a.colsum) ? 1 : 2);
            float4 av_c[NA], av_n[NA], bv_c[NBV], bv_n[NBV];
            float2 bs_c[NBV], bs_n[NBV];
            float4 csum = zero4;                         // column sums of dZ over this CTA's rows (colsum mode)
            int d1[NBV], d2[NBV];                        // gathered row ids of chunk it+1 / it+2 (layer 0)
            auto load_ids = [&](int it, int* d) {
#pragma unroll
                for (int e = 0; e < NBV; ++e) {
                    const int m = c_begin + it * BLOCK_K + ml_b + e * ROWS_B;
                    d[e] = (m < c_end) ? (a.docid ? a.docid[m] : m) : -1;
                }
            };
            auto load_chunk = [&](int it, const int* d, float4* av, float4* bv, float2* bs) {
                const int c0 = c_begin + it * BLOCK_K;
#pragma unroll
                for (int e = 0; e < NA; ++e) {
                    const int m = c0 + ml_a + e * 16;
                    av[e] = (a_ok && m < c_end) ? ld4(a.dZ + (size_t)m * a.N + n_col) : zero4;
                }
#pragma unroll
                for (int e = 0; e < NBV; ++e) {
                    const int m = c0 + ml_b + e * ROWS_B;
                    float4 v = zero4;
                    float2 stv = make_float2(0.f, 0.f);
                    if (d[e] >= 0 && b_kind != 2) {
                        const float* x = a.X + (size_t)d[e] * a.K;
                        stv = a.stats[m];
                        if (b_kind == 0) {
                            v = ld4(x + k_col);
                        } else {
                            // tail of the row: pre-biased so that (v - mean) * rstd gives xhat, 1 (ones column -> db) or 0
                            float t[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int k = k_col + q;
                                t[q] = (k < a.K) ? x[k] : stv.x + (k == a.K ? 1.f / stv.y : 0.f);
                            }
                            v = make_float4(t[0], t[1], t[2], t[3]);
                        }
                    }
                    bv[e] = v;
                    bs[e] = stv;
                }
            };
            if (n_chunks > 0) {
                load_ids(0, d1);
                load_chunk(0, d1, av_c, bv_c, bs_c);
                load_ids(1, d1);
                load_ids(2, d2);
            }
            for (int it = 0; it < n_chunks; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                if (it + 1 < n_chunks) load_chunk(it + 1, d1, av_n, bv_n, bs_n);
#pragma unroll
                for (int e = 0; e < NBV; ++e) d1[e] = d2[e];
                load_ids(it + 3, d2);
                mbar_wait(&empty_bar[s], ph ^ 1u);
                uint8_t* a_hi = smem + s * STAGE_BYTES;
                uint8_t* a_lo = a_hi + A_TILE_BYTES;
                uint8_t* b_hi = a_lo + A_TILE_BYTES;
                uint8_t* b_lo = b_hi + B_TILE_BYTES;
#pragma unroll
                for (int e = 0; e < NA; ++e) {
                    st_split(a_hi, a_lo, swz_mn32(ml_a + e * 16, blk_a, c, 4), av_c[e]);
                    csum.x += av_c[e].x; csum.y += av_c[e].y; csum.z += av_c[e].z; csum.w += av_c[e].w;
                }
#pragma unroll
                for (int e = 0; e < NBV; ++e) {
                    float4 v = bv_c[e];
                    v.x = (v.x - bs_c[e].x) * bs_c[e].y;
                    v.y = (v.y - bs_c[e].x) * bs_c[e].y;
                    v.z = (v.z - bs_c[e].x) * bs_c[e].y;
                    v.w = (v.w - bs_c[e].x) * bs_c[e].y;
                    st_split(b_hi, b_lo, swz_mn32(ml_b + e * ROWS_B, blk_b, c, NB), v);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[s]);
#pragma unroll
                for (int e = 0; e < NA; ++e) av_c[e] = av_n[e];
#pragma unroll
                for (int e = 0; e < NBV; ++e) {
                    bv_c[e] = bv_n[e];
                    bs_c[e] = bs_n[e];
                }
            }
            csum_keep = csum;
        }
    } else if (lane == 0) {
        // =========================== MMA issuer (one thread) ===========================
        constexpr uint32_t idesc =
            make_idesc_tf32(BLOCK_N, KIND == KIND_WGRAD ? 1 : 0, KIND == KIND_WGRAD ? 1 : 0);
        for (int it = 0; it < n_chunks; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES);
            const uint32_t a_lo = a_hi + A_TILE_BYTES;
            const uint32_t b_hi = a_lo + A_TILE_BYTES;
            const uint32_t b_lo = b_hi + B_TILE_BYTES;
            const int rem = c_end - (c_begin + it * BLOCK_K);
            const int nk8 = rem >= BLOCK_K ? 4 : (rem + 7) / 8;
            for (int k8 = 0; k8 < nk8; ++k8) {
                uint64_t da_hi, da_lo, db_hi, db_lo;
                if (KIND != KIND_WGRAD) {
                    // K-major, 128B swizzle: 8-row groups 1024 B apart; one K=8 step = 32 bytes along the row
                    da_hi = make_smem_desc(a_hi + k8 * 32, 16, 1024);
                    da_lo = make_smem_desc(a_lo + k8 * 32, 16, 1024);
                    db_hi = make_smem_desc(b_hi + k8 * 32, 16, 1024);
                    db_lo = make_smem_desc(b_lo + k8 * 32, 16, 1024);
                } else {
                    // MN-major tf32: 128B swizzle with 32B base; 32-element MN blocks 512 B apart (LBO), 4-row K groups
                    // nblk*512 B apart (SBO); one K=8 step = two K groups
                    da_hi = make_smem_desc(a_hi + k8 * 4 * 1024, 512, 4 * 512, kSwizzle128B_Base32B);
                    da_lo = make_smem_desc(a_lo + k8 * 4 * 1024, 512, 4 * 512, kSwizzle128B_Base32B);
                    db_hi = make_smem_desc(b_hi + k8 * NB * 1024, 512, NB * 512, kSwizzle128B_Base32B);
                    db_lo = make_smem_desc(b_lo + k8 * NB * 1024, 512, NB * 512, kSwizzle128B_Base32B);
                }
                const uint32_t acc = (it | k8) != 0 ? 1u : 0u;
                mma_tf32(tmem_base + BLOCK_N, da_lo, db_hi, idesc, acc);
                mma_tf32(tmem_base + BLOCK_N, da_hi, db_lo, idesc, 1u);
                mma_tf32(tmem_base, da_hi, db_hi, idesc, acc);
            }
            mma_commit(&empty_bar[s]);       // ring slot reusable once these MMAs have read it
        }
        mma_commit(accum_bar);               // accumulator complete
        TC_STAMP(3);
    }

    __syncwarp();
    if (warp < MMA_WARP) {
        // =========================== epilogue (all 16 producer warps) ===========================
        // warp w reads TMEM lanes 32*(w%4).. (its rows) and the column blocks cb = w/4, w/4 + 4, ...
        mbar_wait(accum_bar, 0);
        __syncwarp();
        tc_fence_after();
        if (tid == 0) TC_STAMP(4);
        // 1) accumulator (main + correction) -> bias/ELU -> shared-memory tile [128][BLOCK_N + 4] (the operand ring is
        //    free now).  Thread = row in TMEM; the +4 padding keeps the 16-byte row-strided stores conflict free.
        // 2) coalesced copy-out: consecutive lanes write consecutive 16-byte chunks of a row (a thread-per-row store
        //    touches 32 cache lines per instruction and made the epilogue half of the kernel time).
        constexpr int TS = BLOCK_N + 4;                     // tile row stride in floats
        float* stile = reinterpret_cast<float*>(smem);
        // colsum mode: per-warp column sums of dZ [16 warps][32 lanes] float4, behind the tile (the ring is free now)
        constexpr int CS_OFF = (BLOCK_M * TS * 4 + 1023) / 1024 * 1024;
        static_assert(CS_OFF + 16 * 32 * 16 <= STAGES * STAGE_BYTES, "column-sum scratch does not fit the ring");
        const bool do_colsum = KIND == KIND_WGRAD && a.colsum && j0 == 0;
        if (do_colsum) reinterpret_cast<float4*>(smem + CS_OFF)[warp * 32 + lane] = csum_keep;
        const int q4 = warp & 3;
        const int trow = q4 * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
#pragma unroll 1
        for (int cb = warp >> 2; cb < BLOCK_N / 32; cb += 4) {
            float v[32], corr[32];
            tmem_ld32x2(taddr + cb * 32, taddr + BLOCK_N + cb * 32, v, corr);
            float* trow_ptr = stile + (size_t)trow * TS + cb * 32;
#pragma unroll
            for (int q = 0; q < 32; q += 4) {
                float4 o = make_float4(v[q] + corr[q], v[q + 1] + corr[q + 1], v[q + 2] + corr[q + 2],
                                       v[q + 3] + corr[q + 3]);
                if (KIND == KIND_FWD) {
                    const float4 b = *reinterpret_cast<const float4*>(sbias + cb * 32 + q);
                    o.x = elu_fast(o.x + b.x);
                    o.y = elu_fast(o.y + b.y);
                    o.z = elu_fast(o.z + b.z);
                    o.w = elu_fast(o.w + b.w);
                }
                *reinterpret_cast<float4*>(trow_ptr + q) = o;
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");     // the 16 epilogue warps only
        int row_limit, col_limit;
        if (KIND == KIND_FWD) {
            row_limit = a.M; col_limit = a.N;
        } else if (KIND == KIND_DGRAD) {
            row_limit = a.M; col_limit = a.K;
        } else {
            row_limit = a.N; col_limit = a.colsum ? a.K : a.ldo;   // colsum: columns K.. are written below
        }
        if (KIND == KIND_DGRAD && a.fuse_lnbwd) {
            // fused LayerNorm backward + ELU': the tile holds dXhat for FULL rows (BLOCK_N == K).  One warp per row:
            //   dX = rstd * (dXhat - mean_k(dXhat) - xhat * mean_k(dXhat * xhat)),  dZ_prev = dX * ELU'(x)
            for (int r = warp; r < BLOCK_M; r += NPROD / 32) {
                const int grow = i0 + r;
                if (grow >= a.M) break;
                const float* x = a.X + (size_t)grow * a.K;
                const float* g = stile + (size_t)r * TS;
                const float2 st = a.stats[grow];
                float s1 = 0.f, s2 = 0.f;
                for (int k = lane; k < a.K; k += 32) {
                    const float xh = (x[k] - st.x) * st.y;
                    const float d = g[k];
                    s1 += d;
                    s2 = fmaf(d, xh, s2);
                }
                s1 = warp_sum(s1) / (float)a.K;
                s2 = warp_sum(s2) / (float)a.K;
                for (int k = lane; k < a.K; k += 32) {
                    const float xv = x[k];
                    const float xh = (xv - st.x) * st.y;
                    const float dx = st.y * (g[k] - s1 - xh * s2);
                    a.out[(size_t)grow * a.ldo + k] = dx * elu_grad_from_out(xv);
                }
            }
        } else {
            float* out_base = a.out;
            if (KIND == KIND_WGRAD) out_base += (size_t)blockIdx.y * a.N * a.ldo;
            if (do_colsum && tid < BLOCK_M && i0 + tid < a.N) {
                // db partial of row n = i0 + tid: the 16 producer warps covered disjoint contraction rows; lane
                // (n / 32) * 8 + (n % 32) / 4 of every warp holds columns 4 * (n / 4) .. + 3.  Fixed order.
                const float* cs = reinterpret_cast<const float*>(smem + CS_OFF);
                const int src_lane = (tid >> 5) * 8 + ((tid & 31) >> 2);
                float sum = 0.f;
#pragma unroll
                for (int w = 0; w < NPROD / 32; ++w) sum += cs[(w * 32 + src_lane) * 4 + (tid & 3)];
                float* orow = out_base + (size_t)(i0 + tid) * a.ldo;
                orow[a.K] = sum;
                for (int k = a.K + 1; k < a.ldo; ++k) orow[k] = 0.f;
            }
            constexpr int CPR = BLOCK_N / 4;                    // 16-byte chunks per tile row
#pragma unroll 4
            for (int idx = tid; idx < BLOCK_M * CPR; idx += NPROD) {
                const int r = idx / CPR, ch = idx % CPR;
                const int grow = i0 + r, gcol = j0 + ch * 4;
                if (grow < row_limit && gcol < col_limit)         // limits are multiples of 4
                    *reinterpret_cast<float4*>(out_base + (size_t)grow * a.ldo + gcol) =
                        *reinterpret_cast<const float4*>(stile + (size_t)r * TS + ch * 4);
            }
        }
    }
    if (tid == 0) TC_STAMP(5);
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BLOCK_N);
    }
    if (tid == 0) TC_STAMP(6);
}

// Pre-split the weights once per step: forward operand Wf[n][Kpad] = W[n][k] and data-gradient operand
// Wd[k][Npad] = W[n][k] * gamma[k] (transposed), each as (hi, lo) with zero padding to a multiple of 32.
__global__ void __launch_bounds__(256) prep_weights_kernel(PrepTable t) {
    griddep_launch();
    griddep_wait();
    // Output "images": for every 32-wide contraction chunk, [rows x 128 B] in the exact 128B-swizzled shared-memory
    // order, so that a tile (any multiple-of-8 row range of one chunk) is ONE contiguous cp.async.bulk copy.
    //   forward operand   rows = n (N),  contraction = k: value W[n][k]
    //   data-grad operand rows = k (K),  contraction = n: value W[n][k] * gamma[k]
    const int j = blockIdx.y;
    const int K = t.K[j], N = t.N[j], Kpad = t.Kpad[j], Npad = t.Npad[j];
    const float* W = t.W[j];
    const size_t nf = (size_t)N * Kpad;
    const size_t nd = t.wd_hi[j] ? (size_t)K * Npad : 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nf + nd; i += (size_t)gridDim.x * blockDim.x) {
        float w, *hi, *lo;
        size_t o;
        if (i < nf) {
            const int n = (int)(i / Kpad), k = (int)(i % Kpad);      // coalesced read of W along k
            w = k < K ? W[(size_t)n * K + k] : 0.f;
            hi = t.wf_hi[j]; lo = t.wf_lo[j];
            o = (size_t)(k >> 5) * N * 32 + (swz128(n, (k & 31) >> 2) >> 2) + (k & 3);
        } else {
            const size_t q = i - nf;
            const int n = (int)(q / K), k = (int)(q % K);            // n < Npad; coalesced read of W along k
            w = n < N ? W[(size_t)n * K + k] * t.gamma[j][k] : 0.f;
            hi = t.wd_hi[j]; lo = t.wd_lo[j];
            o = (size_t)(n >> 5) * K * 32 + (swz128(k, (n & 31) >> 2) >> 2) + (n & 3);
        }
        float h, l;
        split_tf32(w, h, l);
        hi[o] = h;
        lo[o] = l;
    }
}

template <int KIND, int BLOCK_N>
static cudaError_t launch_one(const TcArgs& a, dim3 grid, cudaStream_t st) {
    constexpr int smem = num_stages(BLOCK_N) * stage_bytes(BLOCK_N) + 1024 + 256 + BLOCK_N * 4;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<KIND, BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    return launch_k(tc_gemm_kernel<KIND, BLOCK_N>, grid, NTHREADS, smem, st, a);
}

template <int KIND>
static cudaError_t launch_kind(const TcArgs& a, int block_n, dim3 grid, cudaStream_t st) {
    if (block_n == 64) return launch_one<KIND, 64>(a, grid, st);
    if (block_n == 128) return launch_one<KIND, 128>(a, grid, st);
    return launch_one<KIND, 256>(a, grid, st);
}

// tile width (dividing `cols`, a multiple of 64) minimising  waves x per-tile cost  (fixed part + width part)
static int pick_block_n(int cols, int row_tiles) {
    const int cands[3] = {256, 128, 64};
    int best = 64, best_cost = 1 << 30;
    for (int i = 0; i < 3; ++i) {
        const int bn = cands[i];
        if (cols % bn != 0) continue;
        const int tiles = row_tiles * (cols / bn);
        const int waves = (tiles + kNumSMs - 1) / kNumSMs;
        const int cost = waves * (4 + bn / 64);
        if (cost < best_cost) {
            best_cost = cost;
            best = bn;
        }
    }
    return best;
}

}  // namespace tc

#ifdef UB200_TC_TIMELINE
extern "C" UB200_API int ub200_tc_timeline(unsigned long long* out64) {
    return (int)cudaMemcpyFromSymbol(out64, tc::g_tc_timeline, sizeof(unsigned long long) * 64);
}
#endif

// ---- entry points used by mlp.cu -------------------------------------------------------------------------
bool tc_layer_ok(int j, int K, int N) {
    return (K % 4 == 0) && (N % 64 == 0) && (j == 0 || K % 64 == 0);
}

int tc_prep(const tc::PrepTable& t, int max_elems, cudaStream_t st) {
    int bx = (max_elems + 256 * 4 - 1) / (256 * 4);
    if (bx > 4 * kNumSMs) bx = 4 * kNumSMs;
    if (bx < 1) bx = 1;
    launch_k(tc::prep_weights_kernel, dim3(bx, t.n), 256, 0, st, t);
    UB_LAUNCH_CHECK("prep_weights_kernel");
    return 0;
}

int tc_forward_layer(const tc::TcArgs& a, cudaStream_t st) {
    const int row_tiles = (a.M + tc::BLOCK_M - 1) / tc::BLOCK_M;
    const int bn = tc::pick_block_n(a.N, row_tiles);
    cudaError_t e = tc::launch_kind<tc::KIND_FWD>(a, bn, dim3(row_tiles, a.N / bn, 1), st);
    count_launch();
    UB_CHECK(e == cudaSuccess, 100, "tc_gemm_kernel<FWD> launch failed: %s", cudaGetErrorString(e));
    return 0;
}

int tc_dgrad_layer(const tc::TcArgs& a, cudaStream_t st) {
    const int row_tiles = (a.M + tc::BLOCK_M - 1) / tc::BLOCK_M;
    const int bn = a.fuse_lnbwd ? a.K : tc::pick_block_n(a.K, row_tiles);   // fused LN-backward needs full rows
    cudaError_t e = tc::launch_kind<tc::KIND_DGRAD>(a, bn, dim3(row_tiles, a.K / bn, 1), st);
    count_launch();
    UB_CHECK(e == cudaSuccess, 100, "tc_gemm_kernel<DGRAD> launch failed: %s", cudaGetErrorString(e));
    return 0;
}

static int wgrad_block_n(int cols) { return cols > 128 ? 256 : (cols > 64 ? 128 : 64); }

// split count of the weight-gradient contraction over the M rows: one wave of CTAs on `sm_budget` SMs (the side
// branches of the backward pass leave the rest to the concurrent data-gradient kernel), each split >= 128 rows
int tc_wgrad_splits(int M, int N, int K, int sm_budget) {
    const int cols = tc_wgrad_colsum(K) ? K : K + 1, bn = wgrad_block_n(cols);
    const int tiles = ((N + tc::BLOCK_M - 1) / tc::BLOCK_M) * ((cols + bn - 1) / bn);
    int s = sm_budget / tiles;
    const int max_s = (M + 127) / 128;
    if (s > max_s) s = max_s;
    return s < 1 ? 1 : s;
}

int tc_wgrad_layer(const tc::TcArgs& a, int splits, cudaStream_t st) {
    const int row_tiles = (a.N + tc::BLOCK_M - 1) / tc::BLOCK_M;
    const int cols = a.colsum ? a.K : a.K + 1;
    const int bn = wgrad_block_n(cols);
    cudaError_t e = tc::launch_kind<tc::KIND_WGRAD>(a, bn, dim3(row_tiles, splits, (cols + bn - 1) / bn), st);
    count_launch();
    UB_CHECK(e == cudaSuccess, 100, "tc_gemm_kernel<WGRAD> launch failed: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace ub200
