// Shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ultra_b200.h"

namespace ub200 {

constexpr int kWarp = 32;
constexpr float kLnEps = 1e-5f;        // nn.LayerNorm default (ultra/ranking_model/DNN.py:46-47)
constexpr int kNumSMs = 148;           // B200

void set_error(const char* fmt, ...);
void count_launch();

#define UB_CHECK(cond, code, ...)            \
    do {                                     \
        if (!(cond)) {                       \
            ub200::set_error(__VA_ARGS__);   \
            return (code);                   \
        }                                    \
    } while (0)

#define UB_LAUNCH_CHECK(name)                                                              \
    do {                                                                                   \
        cudaError_t e__ = cudaGetLastError();                                              \
        ub200::count_launch();                                                             \
        if (e__ != cudaSuccess) {                                                          \
            ub200::set_error("%s launch failed: %s", (name), cudaGetErrorString(e__));     \
            return 100;                                                                    \
        }                                                                                  \
    } while (0)

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------
// Every kernel of the step calls griddep_launch() first (the next kernel of the stream / graph may start being
// scheduled once ALL blocks of this grid have started, i.e. it only takes SMs this grid leaves idle) and
// griddep_wait() after its own on-chip set-up (barrier init, tensor-memory allocation) and BEFORE its first access
// to global memory: the wait returns when every prerequisite grid has completed and flushed, so the data flow is
// exactly that of ordinary stream order while launch latency and prologues overlap the previous kernel's tail.
// Both are no-ops when the kernel was launched without the programmatic attribute.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Programmatic dependent launch (UB200_PDL=1 switches it on; optim.cu).  Every kernel of the library raises
// griddepcontrol.launch_dependents first and touches global memory only behind griddepcontrol.wait, so the next kernel's
// launch latency and on-chip set-up (barrier init, tensor-memory allocation) can overlap the tail of the previous one.
// Measured twice on one B200 (all GPU tests pass with it): -3 % per step for the shortest steps (c1, c2 at L = 10 / 20),
// +1 % at config 2, +4 % at config 5 - no clear win, hence off by default.  Launches of the three big K1 kernels never
// carry the attribute when their grid exceeds one wave (PdlSuppress).
bool pdl_enabled();
int& pdl_suppress();    // thread-local nesting counter: > 0 = the launches of this scope carry no PDL attribute
struct PdlSuppress {
    bool on;
    explicit PdlSuppress(bool cond) : on(cond) { if (on) ++pdl_suppress(); }
    ~PdlSuppress() { if (on) --pdl_suppress(); }
};
// Launch priority of the kernels launched by this host thread from now on (0 = default).  The backward pass gives
// the data-gradient chain (the critical path) the greatest priority and the weight-gradient side branches the
// least, so that when both are pending the block scheduler places the critical kernel's CTAs first.
int& launch_priority();
struct PriorityScope {
    int saved;
    explicit PriorityScope(int p) : saved(launch_priority()) { launch_priority() = p; }
    ~PriorityScope() { launch_priority() = saved; }
};

// <<<grid, block, smem, st>>> with the programmatic-stream-serialization attribute when PDL is enabled
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    unsigned n_at = 0;
    if (pdl_enabled() && pdl_suppress() == 0) {
        at[n_at].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n_at].val.programmaticStreamSerializationAllowed = 1;
        ++n_at;
    }
    if (launch_priority() != 0) {
        at[n_at].id = cudaLaunchAttributePriority;
        at[n_at].val.priority = launch_priority();
        ++n_at;
    }
    cfg.attrs = n_at ? at : nullptr;
    cfg.numAttrs = n_at;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(static_cast<Args&&>(args))...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ELU (alpha = 1): torch uses expm1 on the negative side
__device__ __forceinline__ float elu_f(float z) { return z > 0.f ? z : expm1f(z); }
// derivative of ELU expressed through its OUTPUT y: z>0 -> 1, else exp(z) = y + 1
__device__ __forceinline__ float elu_grad_from_out(float y) { return y > 0.f ? 1.f : y + 1.f; }

// hidden-layer activations of the reference ranker (base_ranking_model.py:63-69: elu, relu, selu, tanh, sigmoid).  The
// tensor-core kernels implement ELU (the reference default); the others run through the fp32 CUDA-core kernels.
// Every derivative is a function of the OUTPUT, so the backward pass needs the stored activations only.
enum { UB200_ACT_ELU = 0, UB200_ACT_RELU = 1, UB200_ACT_SELU = 2, UB200_ACT_TANH = 3, UB200_ACT_SIGMOID = 4 };
constexpr float kSeluAlpha = 1.6732632423543772848170429916717f;     // base_ranking_model.py:13-17
constexpr float kSeluScale = 1.0507009873554804934193349852946f;
__device__ __forceinline__ float act_fwd(float z, int act) {
    switch (act) {
        case UB200_ACT_RELU: return fmaxf(z, 0.f);
        case UB200_ACT_SELU: return kSeluScale * (z >= 0.f ? z : kSeluAlpha * expm1f(z));
        case UB200_ACT_TANH: return tanhf(z);
        case UB200_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
        default: return elu_f(z);
    }
}
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
    switch (act) {
        case UB200_ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case UB200_ACT_SELU: return y >= 0.f ? kSeluScale : y + kSeluScale * kSeluAlpha;
        case UB200_ACT_TANH: return 1.f - y * y;
        case UB200_ACT_SIGMOID: return y * (1.f - y);
        default: return elu_grad_from_out(y);
    }
}

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Deterministic "last block finishes the reduction" ticket. `counter` must be 0 on entry and is reset to 0.
__device__ __forceinline__ bool last_block_ticket(unsigned int* counter, unsigned int nblocks) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == nblocks - 1);
        if (is_last) *counter = 0u;
    }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last;
}

struct LayerDims {
    int n_layers;                   // linear layers including the final ->1
    int K[UB200_MAX_LAYERS];
    int N[UB200_MAX_LAYERS];
    size_t off_g[UB200_MAX_LAYERS], off_b[UB200_MAX_LAYERS], off_w[UB200_MAX_LAYERS], off_c[UB200_MAX_LAYERS];
    size_t n_params;
    int act;                        // UB200_ACT_* of the hidden layers (make_dims: ELU)
};

inline int make_dims(int F, const int* hidden, int n_hidden, LayerDims* d) {
    if (n_hidden < 0 || n_hidden + 1 > UB200_MAX_LAYERS || F <= 0) return 1;
    d->n_layers = n_hidden + 1;
    d->act = UB200_ACT_ELU;
    size_t off = 0;
    int k = F;
    for (int j = 0; j <= n_hidden; ++j) {
        int n = (j == n_hidden) ? 1 : hidden[j];
        if (n <= 0) return 1;
        d->K[j] = k;
        d->N[j] = n;
        d->off_g[j] = off; off += k;
        d->off_b[j] = off; off += k;
        d->off_w[j] = off; off += (size_t)n * k;
        d->off_c[j] = off; off += n;
        k = n;
    }
    d->n_params = off;
    return 0;
}

}  // namespace ub200
