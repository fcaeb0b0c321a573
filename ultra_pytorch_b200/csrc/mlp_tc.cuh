// Host-side entry points of the tensor-core (tcgen05) GEMM path, shared between mlp.cu and mlp_tc.cu.
#pragma once
#include "common.cuh"

namespace ub200 {
namespace tc {
struct TcArgs {
    int M, K, N;
    const float* X;
    const int32_t* docid;
    const float2* stats;
    const float* gamma;
    const float* beta;
    const float* dZ;
    const float* Bhi;            // pre-split, pre-swizzled weight images [chunk][ldb rows x 32]
    const float* Blo;
    int ldb;                     // rows of the image (N for the forward operand, K for the data-gradient operand)
    const float* bias;
    float* out;
    int ldo;
    int rows_per_split;
    // DGRAD with one column tile (BLOCK_N == K): fuse LayerNorm-backward . ELU' into the epilogue and write dZ_{j-1}
    // directly (X = Y_{j-1} is both the LayerNorm input and the ELU output; stats = its row statistics)
    int fuse_lnbwd;
    // WGRAD when K is a multiple of 64: the bias gradient db[n] = sum_m dZ[m, n] is accumulated by the A-operand
    // producers (which stream dZ anyway) instead of a "ones" column of the B operand, which would otherwise open an
    // extra, almost empty column tile (K = 256 -> 257 columns -> two 256-wide tiles) or double the tile width
    int colsum;
};
struct PrepTable {
    int n;
    const float* W[UB200_MAX_LAYERS];
    const float* gamma[UB200_MAX_LAYERS];
    float* wf_hi[UB200_MAX_LAYERS];
    float* wf_lo[UB200_MAX_LAYERS];
    float* wd_hi[UB200_MAX_LAYERS];   // nullptr: not needed
    float* wd_lo[UB200_MAX_LAYERS];
    int K[UB200_MAX_LAYERS], N[UB200_MAX_LAYERS], Kpad[UB200_MAX_LAYERS], Npad[UB200_MAX_LAYERS];
};
// arguments of the fully fused forward kernel (mlp_fused.cu)
struct FusedArgs {
    int M, L, B;                                  // rows = L * B, scores are [B, L]
    int n_hidden;                                 // hidden layers (>= 1)
    int K0;                                       // feature size
    int N[UB200_MAX_LAYERS];                      // hidden sizes
    const float* feats;
    const int32_t* docid;
    const float* gamma[UB200_MAX_LAYERS];         // LayerNorm j (input of linear j), j = 0 .. n_hidden (final)
    const float* beta[UB200_MAX_LAYERS];
    const float* bias[UB200_MAX_LAYERS];          // hidden-layer biases
    const float* wimg_hi[UB200_MAX_LAYERS];       // pre-split weight images (prep_weights_kernel)
    const float* wimg_lo[UB200_MAX_LAYERS];
    const float* w_final;                         // [N_last]
    const float* c_final;                         // [1]
    float* Y[UB200_MAX_LAYERS];                   // activations for the backward pass (write_acts)
    float2* stats[UB200_MAX_LAYERS];              // LayerNorm statistics of every layer input (write_acts)
    float* scores;
    int write_acts;
};
}  // namespace tc

bool fused_forward_ok(int F, const int* N, int n_hidden);
int fused_forward(const tc::FusedArgs& a, cudaStream_t st);
bool tc_layer_ok(int j, int K, int N);
int tc_prep(const tc::PrepTable& t, int max_elems, cudaStream_t st);
int tc_forward_layer(const tc::TcArgs& a, cudaStream_t st);
int tc_dgrad_layer(const tc::TcArgs& a, cudaStream_t st);
int tc_wgrad_splits(int M, int N, int K, int sm_budget = kNumSMs);
int tc_wgrad_layer(const tc::TcArgs& a, int splits, cudaStream_t st);
inline bool tc_wgrad_colsum(int K) { return K % 64 == 0; }
}  // namespace ub200
