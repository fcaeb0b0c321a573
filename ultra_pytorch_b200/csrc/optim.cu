// clip_grad_norm_ + Adagrad / SGD on a flat fp32 parameter buffer (sm_100a).
// Replaces BaseAlgorithm.opt_step (base_algorithm.py:208-226) and DLA.separate_gradient_update (dla.py:141-166).
// ONE launch (grid <= number of SMs, so every block is resident): per-block partial sums of squares, an in-kernel grid
// barrier, then every block adds the partials in the same fixed order (deterministic, identical in every block) and
// applies clip + update.  UB200_OPT_FUSED=0 selects the older two-launch form (norm kernel, then update kernel).
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace ub200 {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

constexpr int kOptBlocks = 2 * kNumSMs;
// workspace: [counter | norm (float at +64B)] [partials kOptBlocks]
static size_t opt_ws_bytes() { return 256 + sizeof(float) * kOptBlocks; }

__device__ __forceinline__ float grad_scale(const float* den, float scale_const) {
    return den ? scale_const / den[0] : scale_const;
}

__global__ void __launch_bounds__(256) grad_norm_kernel(const float* __restrict__ g, size_t n,
                                                         const float* __restrict__ den, float scale_const,
                                                         unsigned int* counter, float* __restrict__ norm_slot,
                                                         float* __restrict__ partials, float* __restrict__ norm_out) {
    griddep_launch();
    griddep_wait();
    __shared__ float red[8];
    const float sc = grad_scale(den, scale_const);
    float s = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = g[i] * sc;
        s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t += red[q];
        partials[blockIdx.x] = t;
    }
    if (last_block_ticket(counter, gridDim.x)) {
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int b = 0; b < (int)gridDim.x; ++b) t += partials[b];
            float nrm = sqrtf(t);
            norm_slot[0] = nrm;
            if (norm_out) norm_out[0] = nrm;
        }
    }
}

__global__ void __launch_bounds__(256) clip_update_kernel(float* __restrict__ p, float* __restrict__ g,
                                                           float* __restrict__ state, size_t n,
                                                           const float* __restrict__ den, float scale_const,
                                                           float max_norm, float lr, int mode,
                                                           const float* __restrict__ norm_slot) {
    griddep_launch();
    griddep_wait();
    float sc = grad_scale(den, scale_const);
    if (max_norm > 0.f) {
        float coef = max_norm / (norm_slot[0] + 1e-6f);     // torch.nn.utils.clip_grad_norm_
        sc *= fminf(coef, 1.f);
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float gi = g[i] * sc;
        g[i] = gi;
        float pi = p[i];
        if (mode == 2) {
            pi -= lr * gi;
        } else {
            float ss = gi * gi;
            if (mode == 0) {
                ss += state[i];
                state[i] = ss;
            }
            pi -= lr * gi / (sqrtf(ss) + 1e-10f);              // torch.optim.Adagrad, eps = 1e-10
        }
        p[i] = pi;
    }
}

// norm + clip + update in one launch.  `bar` = {arrive, depart}: zero on entry and on exit.
__global__ void __launch_bounds__(256) clip_update_fused_kernel(float* __restrict__ p, float* __restrict__ g,
                                                                 float* __restrict__ state, size_t n,
                                                                 const float* __restrict__ den, float scale_const,
                                                                 float max_norm, float lr, int mode,
                                                                 unsigned int* bar, float* __restrict__ partials,
                                                                 float* __restrict__ norm_slot,
                                                                 float* __restrict__ norm_out) {
    griddep_launch();
    griddep_wait();
    __shared__ float red[8];
    __shared__ float s_norm;
    float sc = grad_scale(den, scale_const);
    float s = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = g[i] * sc;
        s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t += red[q];
        partials[blockIdx.x] = t;
        // ---- grid barrier: every block is resident (grid <= SM count), so spinning cannot starve a block ----
        __threadfence();
        atomicAdd(&bar[0], 1u);
        const long long t0 = clock64();
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
            if (seen < gridDim.x && clock64() - t0 > (1ll << 31)) {      // ~1 s: never hang the device ...
                // ... and never update the parameters with a clip coefficient from incomplete partial sums: abort the
                // kernel (the error surfaces at the next synchronisation), as the exchange kernel does (peer.cu)
                printf("ub200 optimizer: grid barrier timed out (block %d saw %u of %u)\n", blockIdx.x, seen, gridDim.x);
                __trap();
            }
        } while (seen < gridDim.x);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        // fixed summation order (lane-strided, then the xor butterfly): bitwise the same value in every block
        float tot = 0.f;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) tot += __ldcg(&partials[b]);
        tot = warp_sum(tot);
        if (threadIdx.x == 0) {
            s_norm = sqrtf(tot);
            // last block to leave re-arms the barrier for the next launch
            if (atomicAdd(&bar[1], 1u) == gridDim.x - 1) {
                bar[0] = 0u;
                __threadfence();
                bar[1] = 0u;
                norm_slot[0] = s_norm;
                if (norm_out) norm_out[0] = s_norm;
            }
        }
    }
    __syncthreads();
    if (max_norm > 0.f) {
        float coef = max_norm / (s_norm + 1e-6f);     // torch.nn.utils.clip_grad_norm_
        sc *= fminf(coef, 1.f);
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float gi = g[i] * sc;
        g[i] = gi;
        float pi = p[i];
        if (mode == 2) {
            pi -= lr * gi;
        } else {
            float ss = gi * gi;
            if (mode == 0) {
                ss += state[i];
                state[i] = ss;
            }
            pi -= lr * gi / (sqrtf(ss) + 1e-10f);              // torch.optim.Adagrad, eps = 1e-10
        }
        p[i] = pi;
    }
}

// Early read-back of a step's scalars: copies n floats into MAPPED PINNED host memory and then bumps a sequence number
// there, so that train() can return the loss as soon as the loss kernel has run while the backward pass and the
// optimizer step of the same batch are still executing (the host spins on the sequence number instead of
// synchronising the stream).  dev_counter counts the launches (CUDA-graph replays included).
__global__ void publish_kernel(const float* __restrict__ src, int n, float* host_dst, unsigned int* host_seq,
                               unsigned int* dev_counter) {
    griddep_launch();
    griddep_wait();
    // launch c writes half (c & 1) of host_dst: a reader that is one launch behind (data-parallel steps return the
    // previous step's loss) never races with the launch in flight
    const unsigned int c = dev_counter[0] + 1u;
    if (threadIdx.x < n) host_dst[(c & 1u) * 32u + threadIdx.x] = src[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        dev_counter[0] = c;
        __threadfence_system();
        *reinterpret_cast<volatile unsigned int*>(host_seq) = c;
    }
}

static int env_flag(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? (v[0] != '0') : dflt;
}
bool pdl_enabled() {
    static const int on = env_flag("UB200_PDL", 0);
    return on != 0;
}
int& pdl_suppress() {
    static thread_local int n = 0;
    return n;
}
int& launch_priority() {
    static thread_local int p = 0;
    return p;
}
static bool opt_fused() {
    static const int on = env_flag("UB200_OPT_FUSED", 1);
    return on != 0;
}

}  // namespace ub200

using namespace ub200;

extern "C" UB200_API const char* ub200_last_error(void) { return g_err; }
extern "C" UB200_API int ub200_abi_version(void) { return 1; }
extern "C" UB200_API unsigned long long ub200_launch_count(void) { return g_launches.load(); }

extern "C" UB200_API int ub200_publish(const float* src, int n, float* host_dst, unsigned int* host_seq,
                                       unsigned int* dev_counter, void* stream) {
    UB_CHECK(src && host_dst && host_seq && dev_counter && n > 0 && n <= 32, 2, "publish: bad arguments");
    launch_k(publish_kernel, 1, 32, 0, static_cast<cudaStream_t>(stream), src, n, host_dst, host_seq, dev_counter);
    UB_LAUNCH_CHECK("publish_kernel");
    return 0;
}

// L2 regularisation of the reference algorithms (hparam l2_loss, e.g. ipw_rank.py:153-157: loss += l2 * sum(p^2) / 2 over
// every ranker parameter): adds l2 * p * f to the UN-normalised gradient buffer, f = den[0] (the normaliser the update
// kernel later divides by) or `factor`, and writes sum(p^2) / 2.  One block: fixed summation order.
__global__ void __launch_bounds__(1024) l2_term_kernel(const float* __restrict__ p, float* __restrict__ g, size_t n, float l2,
                                                       const float* __restrict__ den, float factor,
                                                       float* __restrict__ half_sumsq) {
    griddep_launch();
    griddep_wait();
    __shared__ float red[32];
    const float f = l2 * (den ? den[0] : factor);
    float s = 0.f;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = p[i];
        g[i] = fmaf(f, v, g[i]);
        s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) half_sumsq[0] = 0.5f * t;
    }
}

extern "C" UB200_API int ub200_l2_term(const float* params, float* grads, size_t n, float l2, const float* den, float factor,
                                       float* half_sumsq, void* stream) {
    UB_CHECK(params && grads && half_sumsq && n > 0, 2, "l2_term: null pointer / empty buffer");
    launch_k(l2_term_kernel, 1, 1024, 0, static_cast<cudaStream_t>(stream), params, grads, n, l2, den, factor, half_sumsq);
    UB_LAUNCH_CHECK("l2_term_kernel");
    return 0;
}

extern "C" UB200_API size_t ub200_opt_workspace_bytes(size_t n) {
    (void)n;
    return opt_ws_bytes();
}

extern "C" UB200_API int ub200_clip_update(float* params, float* grads, float* state_sum, size_t n, const float* den,
                                 float scale_const, float max_norm, float lr, int mode, float* norm_out,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    UB_CHECK(params && grads && workspace && n > 0, 2, "clip_update: null pointer / empty buffer");
    UB_CHECK(mode == 1 || mode == 2 || (mode == 0 && state_sum), 1, "clip_update: bad mode %d", mode);
    UB_CHECK(workspace_bytes >= opt_ws_bytes(), 3, "clip_update: workspace too small");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned int* counter = static_cast<unsigned int*>(workspace);
    float* norm_slot = reinterpret_cast<float*>(static_cast<char*>(workspace) + 64);
    float* partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
    int grid = (int)((n + 256 * 4 - 1) / (256 * 4));
    if (grid < 1) grid = 1;
    if ((max_norm > 0.f || norm_out) && opt_fused()) {
        if (grid > kNumSMs) grid = kNumSMs;       // all blocks resident: the in-kernel grid barrier is safe
        launch_k(clip_update_fused_kernel, grid, 256, 0, st, params, grads, state_sum, n, den, scale_const, max_norm,
                 lr, mode, counter + 4, partials, norm_slot, norm_out);
        UB_LAUNCH_CHECK("clip_update_fused_kernel");
        return 0;
    }
    if (grid > kOptBlocks) grid = kOptBlocks;
    if (max_norm > 0.f || norm_out) {
        launch_k(grad_norm_kernel, grid, 256, 0, st, grads, n, den, scale_const, counter, norm_slot, partials,
                 norm_out);
        UB_LAUNCH_CHECK("grad_norm_kernel");
    }
    launch_k(clip_update_kernel, grid, 256, 0, st, params, grads, state_sum, n, den, scale_const, max_norm, lr, mode,
             norm_slot);
    UB_LAUNCH_CHECK("clip_update_kernel");
    return 0;
}
