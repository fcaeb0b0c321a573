// Standalone probe (debug aid, not part of the library): one CTA runs tcgen05.mma.kind::tf32 on operands written to
// shared memory under several layout / descriptor hypotheses and reports which reproduce D = A * B^T exactly.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../ultra_pytorch_b200/csrc/tc_ptx.cuh"
using namespace ub200::tc;

struct Hyp {
    int mn_major;       // 0: K-major operands, 1: MN-major
    int mnblk_stride;   // bytes between 32-element MN blocks (MN-major)
    int kgrp_stride;    // bytes between 8-row K groups
    int lbo, sbo;       // descriptor fields (bytes)
    int use_xor;
    int base32;         // 1: SWIZZLE_128B_BASE32B (4-row K atoms, 32-byte swizzle granularity)
};

template <int N>
__global__ void probe(const float* A, const float* B, float* D, int K, Hyp h) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                 // up to 64 KB
    uint8_t* sb = smem + 65536;         // up to 64 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 131072 / 4; i += blockDim.x) ((float*)smem)[i] = 0.f;
    __syncthreads();
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tslot, N < 32 ? 32 : N);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tmem = tslot;
    // fill operands
    for (int idx = tid; idx < 128 * K; idx += blockDim.x) {
        int mn = idx / K, k = idx % K;
        uint32_t off;
        if (!h.mn_major) off = (mn >> 3) * 1024 + (mn & 7) * 128 + ((((k % 32) / 4) ^ (h.use_xor ? (mn & 7) : 0)) << 4) + (k % 4) * 4 + (k / 32) * 16384;
        else if (!h.base32) off = (mn / 32) * h.mnblk_stride + (k / 8) * h.kgrp_stride + (k % 8) * 128 + ((((mn % 32) / 4) ^ (h.use_xor ? (k % 8) : 0)) << 4) + (mn % 4) * 4;
        else off = (mn / 32) * h.mnblk_stride + (k / 4) * h.kgrp_stride + (k % 4) * 128 + ((((mn % 32) / 8) ^ (h.use_xor ? (k % 4) : 0)) << 5) + (mn % 8) * 4;
        *(float*)(sa + off) = A[mn * K + k];
    }
    for (int idx = tid; idx < N * K; idx += blockDim.x) {
        int mn = idx / K, k = idx % K;
        uint32_t off;
        if (!h.mn_major) off = (mn >> 3) * 1024 + (mn & 7) * 128 + ((((k % 32) / 4) ^ (h.use_xor ? (mn & 7) : 0)) << 4) + (k % 4) * 4 + (k / 32) * (N * 128);
        else if (!h.base32) off = (mn / 32) * h.mnblk_stride + (k / 8) * h.kgrp_stride + (k % 8) * 128 + ((((mn % 32) / 4) ^ (h.use_xor ? (k % 8) : 0)) << 4) + (mn % 4) * 4;
        else off = (mn / 32) * h.mnblk_stride + (k / 4) * h.kgrp_stride + (k % 4) * 128 + ((((mn % 32) / 8) ^ (h.use_xor ? (k % 4) : 0)) << 5) + (mn % 8) * 4;
        *(float*)(sb + off) = B[mn * K + k];
    }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        uint32_t idesc = make_idesc_tf32(N, h.mn_major, h.mn_major);
        for (int k8 = 0; k8 < K / 8; ++k8) {
            uint64_t da, db;
            if (!h.mn_major) {
                da = make_smem_desc(smem_u32(sa) + (k8 / 4) * 16384 + (k8 % 4) * 32, h.lbo, h.sbo);
                db = make_smem_desc(smem_u32(sb) + (k8 / 4) * (N * 128) + (k8 % 4) * 32, h.lbo, h.sbo);
            } else {
                int adv = h.base32 ? 2 * h.kgrp_stride : h.kgrp_stride;
                da = make_smem_desc(smem_u32(sa) + k8 * adv, h.lbo, h.sbo);
                db = make_smem_desc(smem_u32(sb) + k8 * adv, h.lbo, h.sbo);
                if (h.base32) {   // layout type 1 instead of 2
                    da = (da & ~((uint64_t)7 << 61)) | ((uint64_t)1 << 61);
                    db = (db & ~((uint64_t)7 << 61)) | ((uint64_t)1 << 61);
                }
            }
            mma_tf32(tmem, da, db, idesc, k8 ? 1u : 0u);
        }
        mma_commit(&bar);
    }
    __syncwarp();
    if (warp < 4) {
        mbar_wait(&bar, 0);
        __syncwarp();
        tc_fence_after();
        for (int cb = 0; cb < N / 32; ++cb) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + cb * 32, v);
            for (int q = 0; q < 32; ++q) D[(warp * 32 + (tid & 31)) * N + cb * 32 + q] = v[q];
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, N < 32 ? 32 : N); }
}

int main() {
    const int N = 64, K = 16;
    std::vector<float> A(128 * K), B(N * K), D(128 * N), R(128 * N);
    srand(1);
    for (auto& x : A) x = (float)(rand() % 7 - 3);
    for (auto& x : B) x = (float)(rand() % 7 - 3);
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k]; R[m * N + n] = s; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 133120);
    // A has 4 MN blocks, B has 2: a single mnblk/kgrp stride pair must serve both -> use strides valid for 4 blocks
    Hyp hyps[] = {
        {0, 0, 0, 16, 1024, 1, 0},              // K-major reference (known good)
        {1, 2048, 512, 2048, 512, 1, 1},        // B0: K groups inner (K=16 -> 4 groups of 4), LBO = MN-block stride
        {1, 2048, 512, 512, 2048, 1, 1},        // B1: fields swapped
        {1, 512, 2048, 512, 2048, 1, 1},        // B2: MN blocks inner (4 blocks), LBO = MN-block stride
        {1, 512, 2048, 2048, 512, 1, 1},        // B3: fields swapped
        {1, 2048, 512, 2048, 512, 0, 1},        // B4: no xor
        {1, 512, 2048, 512, 2048, 0, 1},        // B5: no xor
    };
    int nh = sizeof(hyps) / sizeof(hyps[0]);
    for (int i = 0; i < nh; ++i) {
        cudaMemset(dD, 0, D.size() * 4);
        probe<N><<<1, 128, 133120>>>(dA, dB, dD, K, hyps[i]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("hyp %d: CUDA error %s\n", i, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0; for (size_t q = 0; q < D.size(); ++q) bad += (D[q] != R[q]);
        printf("hyp %d (mn_major=%d base32=%d blk=%d kgrp=%d lbo=%d sbo=%d xor=%d): %d / %zu mismatches; D[0..3]=%g %g %g %g ref %g %g %g %g\n", i,
               hyps[i].mn_major, hyps[i].base32, hyps[i].mnblk_stride, hyps[i].kgrp_stride, hyps[i].lbo, hyps[i].sbo, hyps[i].use_xor, bad, D.size(),
               D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
    }
    return 0;
}
