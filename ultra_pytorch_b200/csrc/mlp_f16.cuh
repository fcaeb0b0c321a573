// Host-side entry points of the fp16-split tensor-core path (mlp_f16.cu), shared with mlp.cu.
//
// Every fp32 operand value x of a dense contraction is split as x = hi + lo with hi, lo fp16 (22 significant bits, the
// same as the 3xTF32 split) and  hi*hi + lo*hi + hi*lo  is issued as three tcgen05.mma.kind::f16 (K = 16 per
// instruction): twice the tensor-core rate and half the shared-memory bytes of the TF32 form.  Range is handled by
// exact power-of-two scales: weights x 2^8 (|w| ~ 1/sqrt(K) would put the lo part into fp16 subnormals), gradients by
// a per-row (data gradient) or per-tensor (weight gradient) power of two derived from the running max.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace ub200 {
namespace f16 {

constexpr int MAXF = 4;                 // hidden layers one fused launch can chain

struct PrepArgs {
    int n;                                      // hidden layers
    const float* W[UB200_MAX_LAYERS];
    const float* gamma[UB200_MAX_LAYERS];
    const float* beta[UB200_MAX_LAYERS];
    const float* bias[UB200_MAX_LAYERS];
    uint16_t* wf[UB200_MAX_LAYERS];             // forward operand image      [K/64 chunks][hi|lo][N rows][64 k]  of 2^8 W gamma
    uint16_t* wd[UB200_MAX_LAYERS];             // data-gradient operand image [N/64 chunks][hi|lo][K rows][64 n]  (nullptr: skip)
    float* bias2[UB200_MAX_LAYERS];             // b + W beta
    int K[UB200_MAX_LAYERS], N[UB200_MAX_LAYERS];
    const float* gF;                            // final layer (N = 1): LayerNorm weight / bias, weight row, bias
    const float* bF;
    const float* wF;
    const float* cF;
    int KF;
    float* wf2;                                 // [KF] gamma_F * w_F
    float* cf2;                                 // [1]  c_F + beta_F . w_F
    unsigned int* dzmax;                        // [UB200_MAX_LAYERS] running maxima of this step's backward pass: reset here
    // operand images of the weight-gradient kernel (see ImgInfo below): before the maxima are reset, the power-of-two
    // scale of this step's dZ_j images is derived from the maxima the PREVIOUS step left behind; img_bad is cleared
    float* wscale;                              // [UB200_MAX_LAYERS] (0 = no history: no image this step); nullptr: off
    unsigned int* img_bad;                      // [UB200_MAX_LAYERS]
};

struct FwdArgs {
    int M, L, B;
    int nl;                                     // hidden layers chained in this launch (>= 1)
    int has_final;                              // fuse the final LayerNorm -> Linear(1), write scores [B, L]
    int K0;                                     // width of the first layer's input
    int N[MAXF];                                // layer widths
    int bn0;                                    // column tile of layer 0 (== N[0] unless nl == 1 && !has_final: gridDim.y tiles)
    int dual[MAXF];                             // separate correction accumulator (long contractions)
    const float* X;                             // first layer input rows [*, K0]
    const int32_t* docid;                       // row gather of the first layer (nullptr: identity)
    const uint16_t* wimg[MAXF];
    const float* bias2[MAXF];
    const float* wf2;
    const float* cf2;
    float* Y[MAXF];                             // post-ELU activations (write_acts)
    float2* stats[MAXF + 1];                    // (mean, rstd) of the input rows of layer q; [nl] = of the final layer's input
    float* scores;
    int write_acts;
    // TMA descriptors of Y[q] viewed as [M rows, N_q cols] fp32 with a 32-column x 128-row box and the 128B swizzle: the
    // epilogue stages 64-column chunks in shared memory and one thread issues the (fully coalesced) tensor stores
    CUtensorMap ymap[MAXF];
    // training: every A-operand stage this launch builds (normalised input of layer q, fp16 hi | lo, 128 rows x 64
    // columns, 128B swizzle: 32 KB) is also copied to ximg[q] + (row_tile * ceil(K_q / 64) + chunk) * 32 KB by a bulk
    // shared->global copy: the K-major [m][k] tile IS the MN-major B operand of the weight gradient, which then gets its
    // operands from the TMA engine instead of re-reading and re-converting fp32 activations (nullptr: no image)
    uint16_t* ximg[MAXF];
};
// 2-D fp32 tensor map [rows, cols] (row-major), box 32 cols x 128 rows, SWIZZLE_128B
int make_tmap_f32(CUtensorMap* m, const float* base, size_t rows, size_t cols);

struct BwdArgs {
    int M, L, B;
    int nl;                                     // hidden layers chained (dZ of layers nl-1 .. 0 are produced)
    int KF;                                     // == N[nl-1]
    int N[MAXF];                                // widths of hidden layers first .. first+nl-1
    const float* dscores;                       // [B, L]
    const float* wf2;                           // gamma_F * w_F [KF]
    const float* Y[MAXF];                       // post-ELU activations of the chained layers
    const float2* stats[MAXF + 1];              // stats[q] = of layer q's INPUT rows (q >= 1: rows of Y[q-1]); [nl] = final input
    const uint16_t* wd[MAXF];                   // data-gradient images of layers 1 .. nl-1 (index q)
    float* dZ[MAXF];                            // out: dZ_q [M, N_q] fp32
    unsigned int* dzmax[MAXF];                  // out: running max |dZ_q| (float bits, atomicMax) for the weight-gradient scale
    CUtensorMap ymap[MAXF];                     // Y[q] / dZ[q] as [M, N_q] fp32, 32-column x 128-row boxes, 128B swizzle
    CUtensorMap dzmap[MAXF];
    // weight-gradient operand images of dZ_q (same tile format as FwdArgs::ximg; (row_tile * N_q / 64 + chunk) * 32 KB).
    // They carry ONE power-of-two scale per layer, wscale[q], derived from the previous step's max |dZ_q| (this step's
    // is only known when the kernel is over).  The data-gradient GEMM uses the same tiles; a row whose values would
    // leave the fp16 range under that scale falls back to its own row scale and raises img_bad[q], which sends the
    // weight gradient of that layer down the fp32 path for this step.  nullptr / wscale == 0: no image.
    uint16_t* dzimg[MAXF];
    const float* wscale;                        // [>= nl] (device)
    unsigned int* img_bad;                      // [>= nl] (device)
};

size_t img_bytes(int M, int cols);                    // operand image of an [M, cols] activation / gradient matrix
size_t prep_bytes_wf(int K, int N);
size_t prep_bytes_wd(int K, int N);
int prep(const PrepArgs& a, cudaStream_t st);
bool fwd_shape_ok(int K0, const int* N, int nl);      // can fwd() chain these layers
int fwd(const FwdArgs& a, cudaStream_t st);
bool bwd_shape_ok(const int* N, int nl);
int bwd(const BwdArgs& a, cudaStream_t st);
int bwd_grid(int M);

}  // namespace f16
}  // namespace ub200

// ---- weight gradients of ALL hidden layers in one launch (mlp_f16.cu: wgrad16_kernel) --------------------------------
namespace ub200 {
namespace f16 {

struct WgLayer {
    const float* dZ;              // [M, N]
    const float* X;               // layer input rows [*, K] (features for layer 0, else Y_{j-1})
    const int32_t* docid;         // row gather (layer 0) or nullptr
    const float2* stats;          // (mean, rstd) of the input rows [M]
    const unsigned int* dzmax;    // max |dZ| (float bits) -> power-of-two operand scale
    float* out;                   // partial planes [splits][N][ldp]; column K holds the bias-gradient partial
    int N, K, ldp;
    int bn;                       // column tile of the B operand in shared memory (64 / 128 / 192 / 256)
    int m_tiles, col_tiles, splits, rows_per_split;
    int cta_begin;                // first CTA of this layer in the grid
    // operand images written by the forward / backward kernels (nullptr: convert from fp32 in this kernel)
    const uint16_t* ximg;         // [M / 128 tiles][ceil(K / 64) chunks][hi | lo][128 x 64]
    const uint16_t* dzimg;        // [M / 128 tiles][N / 64 chunks][hi | lo][128 x 64], scaled by *wscale
    const float* wscale;
    const unsigned int* img_bad;  // != 0: this step's dZ image is not usable
};
struct WgArgs {
    int n, M, total_ctas;
    WgLayer l[UB200_MAX_LAYERS];
};

// fills bn / tiles / splits / cta_begin of every layer for a budget of `sm_budget` CTAs (one wave); rows per split are
// capped so that one accumulator never sees more than 2048 contraction rows (accumulation error, partial-plane traffic)
void wgrad_plan(WgArgs* a, int sm_budget);
int wgrad(const WgArgs& a, cudaStream_t st);

}  // namespace f16
}  // namespace ub200
