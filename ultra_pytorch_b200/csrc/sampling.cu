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

}  // namespace ub200

using namespace ub200;

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
