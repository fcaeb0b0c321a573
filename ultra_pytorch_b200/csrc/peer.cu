// C1 - data-parallel gradient exchange over NVLink peer memory (sm_100a).
//
// One kernel, launched by every rank on its own GPU inside the step's CUDA graph, replaces the NCCL all-reduce of the
// flat buffer [DNN grads | loss normalisers | EM / DenoisingNet partials] (310 KB at config 2, 2.1 MB
// at config 3 - a latency-bound message):
//
//   A. handshake: every rank stores "my buffer is final" into every peer's flag array (st.release.sys over NVLink)
//      and waits for the same flag from every peer;
//   B. one-shot reduce: every rank reads its slice-interleaved share of ALL peers' buffers straight from peer memory
//      (ld.global.cv, 16 bytes per thread) and adds them in rank order 0..W-1 - the same order on every rank, so the
//      replicas stay bitwise identical - into a local scratch buffer;
//   C. handshake: "I am done reading", then wait until every peer is done reading MY buffer;
//   D. the sums are copied back into the rank's own buffer, where the optimizer kernel that follows finds them.
//
// Buffers and flag arrays live in symmetric memory (torch.distributed._symmetric_memory: the same allocation mapped
// into every rank's address space).  No host involvement, no second stream; every spin is bounded (trap after ~10 s
// instead of hanging the device).
#include "common.cuh"

namespace ub200 {

constexpr int kMaxPeers = 16;

struct PeerTable {
    const float* buf[kMaxPeers];       // rank r's flat gradient buffer, mapped into this rank's address space
    unsigned int* flags[kMaxPeers];    // rank r's flag array [2 * W]: [0, W) "buffer final", [W, 2W) "done reading"
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_cv4(const float* p) {
    float4 v;
    asm volatile("ld.global.cv.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_cv1(const float* p) {
    float v;
    asm volatile("ld.global.cv.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

constexpr long long kSpinLimit = 20000000000ll;       // ~10 s of SM clocks

// wait until *flag (written by a peer) reaches `want`; flags only grow
__device__ __forceinline__ void wait_flag_sys(const unsigned int* flag, unsigned int want) {
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(flag) - want) < 0) {
        if (clock64() - t0 > kSpinLimit) __trap();
    }
}

// ctl (local, zero on first use): [0] sequence number of this launch, [1] arrive, [2] depart of the grid barrier
__global__ void __launch_bounds__(256) peer_allreduce_kernel(PeerTable peers, int rank, int W, float* __restrict__ own,
                                                              float* __restrict__ scratch, size_t n,
                                                              unsigned int* ctl) {
    griddep_launch();
    griddep_wait();
    const unsigned int seq = ld_acquire_gpu(&ctl[0]);       // bumped by the last block of the previous launch
    const unsigned int ready = 2u * seq + 1u, done = 2u * seq + 2u;
    unsigned int* my_flags = peers.flags[rank];

    // ---- A: my buffer is final (all kernels that wrote it completed before this kernel started) ----
    if (blockIdx.x == 0 && threadIdx.x < W) {
        __threadfence_system();
        st_release_sys(peers.flags[threadIdx.x] + rank, ready);
    }
    if (threadIdx.x < W) wait_flag_sys(my_flags + threadIdx.x, ready);
    __syncthreads();

    // ---- B: one-shot reduce in rank order ----
    const size_t n4 = n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 s = ld_cv4(peers.buf[0] + 4 * i);
        for (int r = 1; r < W; ++r) {
            const float4 v = ld_cv4(peers.buf[r] + 4 * i);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        *reinterpret_cast<float4*>(scratch + 4 * i) = s;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - 4 * n4)) {
        const size_t i = 4 * n4 + threadIdx.x;
        float s = ld_cv1(peers.buf[0] + i);
        for (int r = 1; r < W; ++r) s += ld_cv1(peers.buf[r] + i);
        scratch[i] = s;
    }

    // ---- grid barrier: every block of this rank is done reading the peers ----
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&ctl[1], 1u);
        const long long t0 = clock64();
        while (ld_acquire_gpu(&ctl[1]) < gridDim.x) {
            if (clock64() - t0 > kSpinLimit) __trap();
        }
    }
    __syncthreads();

    // ---- C: tell every peer, wait until every peer is done reading MY buffer ----
    if (blockIdx.x == 0 && threadIdx.x < W) st_release_sys(peers.flags[threadIdx.x] + W + rank, done);
    if (threadIdx.x < W) wait_flag_sys(my_flags + W + threadIdx.x, done);
    __syncthreads();

    // ---- D: the sums replace my partials ----
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        *reinterpret_cast<float4*>(own + 4 * i) = *reinterpret_cast<const float4*>(scratch + 4 * i);
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - 4 * n4)) own[4 * n4 + threadIdx.x] = scratch[4 * n4 + threadIdx.x];

    // last block to leave re-arms the grid barrier and advances the sequence number
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&ctl[2], 1u) == gridDim.x - 1) {
            ctl[1] = 0u;
            ctl[2] = 0u;
            __threadfence();
            ctl[0] = seq + 1u;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused exchange + optimizer: ONE kernel = all-reduce(SUM) of the flat buffer over NVLink + clip_grad_norm_ + Adagrad.
//
// Push model, one handshake: block b of every rank STORES its slice of the rank's buffer into every peer's inbox
// (posted NVLink writes, no round trip), publishes a per-(sender, block) flag with st.release.sys, waits for the same
// flag of every peer, and adds own + inboxes in rank order 0..W-1 (bitwise identical sums on every rank).  Inboxes are
// double buffered by step parity: a rank can run at most one step ahead of a peer (it needs the peer's flag of the
// current step), so the slot it overwrites two steps later has been consumed - no "done reading" handshake.  Only
// block b of the peers has to have arrived for block b to proceed.  Then, exactly as clip_update_fused_kernel:
// per-block sum of squares, in-kernel grid barrier, fixed-order norm, clip, Adagrad / SGD - on the summed gradient.
// Afterwards own[0, n_params) holds the clipped summed gradient and own[n_params, n) the summed trailing floats
// (loss normalisers, EM / DenoisingNet partials).
// ---------------------------------------------------------------------------------------------------------------
struct PushTable {
    float* inbox[kMaxPeers];           // rank r's inbox [2][W][n rounded up to a multiple of 4]
    unsigned int* flags[kMaxPeers];    // rank r's flags [W][kNumSMs]
};

// ctl: [0] seq, [1] arrive, [2] depart, floats at +64 B: partials[kNumSMs]
__global__ void __launch_bounds__(256) dp_reduce_update_kernel(PushTable peers, int rank, int W, float* __restrict__ own,
                                                                size_t n, float* __restrict__ p,
                                                                float* __restrict__ state, size_t n_params,
                                                                long long den_index, float scale_const, float max_norm,
                                                                float lr, int mode, float* __restrict__ norm_out,
                                                                unsigned int* ctl, long long pub_index, int pub_n,
                                                                float* host_dst, unsigned int* host_seq,
                                                                unsigned int* dev_counter) {
    griddep_launch();
    griddep_wait();
    __shared__ float red[8];
    __shared__ float s_norm;
    float* partials = reinterpret_cast<float*>(ctl + 16);
    const unsigned int seq = ld_acquire_gpu(&ctl[0]);
    const unsigned int want = seq + 1u;
    const size_t par = seq & 1u;
    const size_t ns = (n + 3) / 4 * 4;                    // inbox stride (keeps every slot 16-byte aligned)
    // contiguous slice of this block, a multiple of 4 floats (the last block takes the ragged tail)
    const size_t per = ((n + gridDim.x - 1) / gridDim.x + 3) / 4 * 4;
    const size_t lo = (size_t)blockIdx.x * per < n ? (size_t)blockIdx.x * per : n;
    const size_t hi = lo + per < n ? lo + per : n;

    // ---- push my slice into every peer's inbox, then publish the flag of (me, this block) ----
    for (int q = 1; q < W; ++q) {
        const int r = (rank + q) % W;                     // start with the next rank: spreads the NVLink traffic
        float* dst = peers.inbox[r] + (par * W + rank) * ns;
        const size_t n4 = (hi - lo) / 4;
        for (size_t i = threadIdx.x; i < n4; i += blockDim.x)
            *reinterpret_cast<float4*>(dst + lo + 4 * i) = *reinterpret_cast<const float4*>(own + lo + 4 * i);
        for (size_t i = lo + 4 * n4 + threadIdx.x; i < hi; i += blockDim.x) dst[i] = own[i];
    }
    __syncthreads();
    if (threadIdx.x < W && (int)threadIdx.x != rank) {
        __threadfence_system();
        st_release_sys(peers.flags[threadIdx.x] + (size_t)rank * kNumSMs + blockIdx.x, want);
    }
    // ---- wait for block b of every peer ----
    if (threadIdx.x < W && (int)threadIdx.x != rank)
        wait_flag_sys(peers.flags[rank] + (size_t)threadIdx.x * kNumSMs + blockIdx.x, want);
    __syncthreads();

    // ---- reduce in rank order; the summed gradient replaces my partial ----
    const float* inbox = peers.inbox[rank] + par * W * ns;
    float ss = 0.f;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        float g = 0.f;
        for (int r = 0; r < W; ++r) g += (r == rank) ? own[i] : __ldcg(inbox + (size_t)r * ns + i);
        own[i] = g;
        if (i < n_params) ss = fmaf(g, g, ss);
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t += red[q];
        partials[blockIdx.x] = t;
        __threadfence();
        atomicAdd(&ctl[1], 1u);
        const long long t0 = clock64();
        while (ld_acquire_gpu(&ctl[1]) < gridDim.x) {
            if (clock64() - t0 > kSpinLimit) __trap();
        }
    }
    __syncthreads();
    float sc = scale_const;
    if (den_index >= 0) sc = scale_const / __ldcg(own + den_index);      // the SUMMED normaliser
    if (threadIdx.x < 32) {
        float tot = 0.f;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) tot += __ldcg(&partials[b]);
        tot = warp_sum(tot);
        if (threadIdx.x == 0) {
            s_norm = sqrtf(tot) * fabsf(sc);                              // || g * sc ||
            if (atomicAdd(&ctl[2], 1u) == gridDim.x - 1) {
                ctl[1] = 0u;
                ctl[2] = 0u;
                __threadfence();
                ctl[0] = seq + 1u;
                if (norm_out) norm_out[0] = s_norm;
            }
        }
    }
    __syncthreads();
    if (pub_n > 0 && blockIdx.x == 0) {
        // the step's summed scalars (loss normalisers in the trailing floats: every block's slice is final behind the
        // barrier) go to mapped pinned host memory exactly as publish_kernel (optim.cu) writes them - train() of the
        // NEXT step reads them, so no separate kernel sits at the end of a data-parallel step
        const unsigned int c = dev_counter[0] + 1u;
        if ((int)threadIdx.x < pub_n) host_dst[(c & 1u) * 32u + threadIdx.x] = __ldcg(own + pub_index + threadIdx.x);
        __syncthreads();
        if (threadIdx.x == 0) {
            dev_counter[0] = c;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned int*>(host_seq) = c;
        }
    }
    if (max_norm > 0.f) sc *= fminf(max_norm / (s_norm + 1e-6f), 1.f);   // torch.nn.utils.clip_grad_norm_
    const size_t hip = hi < n_params ? hi : n_params;
    for (size_t i = lo + threadIdx.x; i < hip; i += blockDim.x) {
        const float gi = own[i] * sc;
        own[i] = gi;
        float pi = p[i];
        if (mode == 2) {
            pi -= lr * gi;
        } else {
            float s2 = gi * gi;
            if (mode == 0) {
                s2 += state[i];
                state[i] = s2;
            }
            pi -= lr * gi / (sqrtf(s2) + 1e-10f);                          // torch.optim.Adagrad, eps = 1e-10
        }
        p[i] = pi;
    }
}

}  // namespace ub200

using namespace ub200;

extern "C" UB200_API size_t ub200_dp_inbox_bytes(int world, size_t n) {
    return sizeof(float) * 2 * (size_t)(world > 0 ? world : 0) * ((n + 3) / 4 * 4);
}
extern "C" UB200_API size_t ub200_dp_flag_bytes(int world) {
    return sizeof(unsigned int) * (size_t)(world > 0 ? world : 0) * kNumSMs;
}
extern "C" UB200_API size_t ub200_dp_ctl_bytes(void) { return 64 + sizeof(float) * kNumSMs + 64; }

static int dp_reduce_update_impl(float* own, size_t n, const void* const* peer_inbox, const void* const* peer_flags,
                                 int rank, int world, float* params, float* state_sum, size_t n_params,
                                 long long den_index, float scale_const, float max_norm, float lr, int mode,
                                 float* norm_out, void* ctl, void* stream, long long pub_index, int pub_n,
                                 float* host_dst, unsigned int* host_seq, unsigned int* dev_counter) {
    UB_CHECK(pub_n == 0 || (pub_n > 0 && pub_n <= 32 && pub_index >= 0 && pub_index + pub_n <= (long long)n && host_dst &&
                            host_seq && dev_counter),
             1, "dp_reduce_update: bad publish arguments");
    UB_CHECK(own && peer_inbox && peer_flags && params && ctl && n > 0 && n_params <= n, 2,
             "dp_reduce_update: null pointer / bad sizes");
    UB_CHECK(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, 1,
             "dp_reduce_update: bad rank %d / world %d", rank, world);
    UB_CHECK(mode == 1 || mode == 2 || (mode == 0 && state_sum), 1, "dp_reduce_update: bad mode %d", mode);
    UB_CHECK(den_index < (long long)n, 1, "dp_reduce_update: normaliser index outside the buffer");
    PushTable t;
    memset(&t, 0, sizeof(t));
    for (int r = 0; r < world; ++r) {
        UB_CHECK(peer_inbox[r] && peer_flags[r], 2, "dp_reduce_update: null peer pointer (rank %d)", r);
        t.inbox[r] = static_cast<float*>(const_cast<void*>(peer_inbox[r]));
        t.flags[r] = static_cast<unsigned int*>(const_cast<void*>(peer_flags[r]));
    }
    UB_CHECK((reinterpret_cast<uintptr_t>(own) & 15) == 0, 1, "dp_reduce_update: unaligned buffer");
    for (int r = 0; r < world; ++r)
        UB_CHECK((reinterpret_cast<uintptr_t>(peer_inbox[r]) & 15) == 0, 1, "dp_reduce_update: unaligned inbox");
    int grid = (int)((n + 1023) / 1024);
    if (grid > kNumSMs) grid = kNumSMs;      // every block resident; the same n gives the same grid on every rank
    if (grid < 1) grid = 1;
    launch_k(dp_reduce_update_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), t, rank, world, own, n, params,
             state_sum, n_params, den_index, scale_const, max_norm, lr, mode, norm_out, static_cast<unsigned int*>(ctl),
             pub_index, pub_n, host_dst, host_seq, dev_counter);
    UB_LAUNCH_CHECK("dp_reduce_update_kernel");
    return 0;
}

extern "C" UB200_API int ub200_dp_reduce_update(float* own, size_t n, const void* const* peer_inbox,
                                                const void* const* peer_flags, int rank, int world, float* params,
                                                float* state_sum, size_t n_params, long long den_index,
                                                float scale_const, float max_norm, float lr, int mode,
                                                float* norm_out, void* ctl, void* stream) {
    return dp_reduce_update_impl(own, n, peer_inbox, peer_flags, rank, world, params, state_sum, n_params, den_index,
                                 scale_const, max_norm, lr, mode, norm_out, ctl, stream, 0, 0, nullptr, nullptr, nullptr);
}

// the same + ub200_publish of own[pub_index, pub_index + pub_n) (the summed scalars of the step) from inside the kernel
extern "C" UB200_API int ub200_dp_reduce_update_publish(float* own, size_t n, const void* const* peer_inbox,
                                                        const void* const* peer_flags, int rank, int world,
                                                        float* params, float* state_sum, size_t n_params,
                                                        long long den_index, float scale_const, float max_norm, float lr,
                                                        int mode, float* norm_out, void* ctl, void* stream,
                                                        long long pub_index, int pub_n, float* host_dst,
                                                        unsigned int* host_seq, unsigned int* dev_counter) {
    return dp_reduce_update_impl(own, n, peer_inbox, peer_flags, rank, world, params, state_sum, n_params, den_index,
                                 scale_const, max_norm, lr, mode, norm_out, ctl, stream, pub_index, pub_n, host_dst,
                                 host_seq, dev_counter);
}

extern "C" UB200_API size_t ub200_peer_ctl_bytes(void) { return 256; }
extern "C" UB200_API size_t ub200_peer_flag_bytes(int world) { return sizeof(unsigned int) * 2 * (size_t)(world > 0 ? world : 0); }

extern "C" UB200_API int ub200_peer_allreduce(const void* const* peer_bufs, const void* const* peer_flags, int rank,
                                              int world, float* scratch, size_t n, void* ctl, void* stream) {
    UB_CHECK(peer_bufs && peer_flags && scratch && ctl && n > 0, 2, "peer_allreduce: null pointer / empty buffer");
    UB_CHECK(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, 1, "peer_allreduce: bad rank %d / world %d",
             rank, world);
    PeerTable t;
    memset(&t, 0, sizeof(t));
    for (int r = 0; r < world; ++r) {
        UB_CHECK(peer_bufs[r] && peer_flags[r], 2, "peer_allreduce: null peer pointer (rank %d)", r);
        UB_CHECK((reinterpret_cast<uintptr_t>(peer_bufs[r]) & 15) == 0, 1, "peer_allreduce: buffer not 16-byte aligned");
        t.buf[r] = static_cast<const float*>(peer_bufs[r]);
        t.flags[r] = static_cast<unsigned int*>(const_cast<void*>(peer_flags[r]));
    }
    int grid = (int)((n / 4 + 255) / 256);
    if (grid > kNumSMs) grid = kNumSMs;      // every block resident: the in-kernel grid barrier cannot starve
    if (grid < 1) grid = 1;
    float* own = const_cast<float*>(t.buf[rank]);
    launch_k(peer_allreduce_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), t, rank, world, own, scratch, n,
             static_cast<unsigned int*>(ctl));
    UB_LAUNCH_CHECK("peer_allreduce_kernel");
    return 0;
}
