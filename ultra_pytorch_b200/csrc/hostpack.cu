// Host-side packing of one input_feed into the pinned staging buffer (multi-threaded, OpenMP).
// Replaces the numpy work of BaseAlgorithm.create_input_feed / get_ranking_scores on the host side of the boundary
// (base_algorithm.py:148-152 concat + np.take, :176-186 label transpose / docid conversion, DNN.py:72-73 f64 -> f32).
#include <omp.h>

#include "common.cuh"

using namespace ub200;

extern "C" UB200_API size_t ub200_feed_bytes(int n_docs, int F, int L, int B) {
    const size_t off_f = align_up((size_t)8 * L * B, 256);
    return off_f + sizeof(float) * (size_t)(n_docs + 1) * F;
}

extern "C" UB200_API int ub200_pack_feed_host(const double* feats, int n_docs, int F, const float* const* docid_cols,
                                              const float* const* label_cols, int L, int B, void* dst,
                                              size_t dst_bytes, int n_threads) {
    UB_CHECK(dst && docid_cols && label_cols && (feats || n_docs == 0), 2, "pack_feed_host: null pointer");
    UB_CHECK(L > 0 && B > 0 && F > 0 && n_docs >= 0, 1, "pack_feed_host: bad sizes");
    const size_t need = ub200_feed_bytes(n_docs, F, L, B);
    UB_CHECK(dst_bytes >= need, 3, "pack_feed_host: destination too small (%zu < %zu)", dst_bytes, need);
    char* base = static_cast<char*>(dst);
    int32_t* docid = reinterpret_cast<int32_t*>(base);                               // [L, B] position-major
    float* labels = reinterpret_cast<float*>(base + (size_t)4 * L * B);              // [B, L]
    float* f32 = reinterpret_cast<float*>(base + align_up((size_t)8 * L * B, 256));   // [n_docs + 1, F]
    if (n_threads < 1) n_threads = 1;
    const long long nf = (long long)n_docs * F;
    const long long blk = 16384;
    const long long nblk = (nf + blk - 1) / blk;
#pragma omp parallel num_threads(n_threads)
    {
#pragma omp for schedule(static) nowait
        for (long long b = 0; b < nblk; ++b) {
            const long long lo = b * blk, hi = lo + blk < nf ? lo + blk : nf;
            for (long long i = lo; i < hi; ++i) f32[i] = (float)feats[i];
        }
#pragma omp for schedule(static) nowait
        for (int l = 0; l < L; ++l) {
            const float* d = docid_cols[l];
            const float* y = label_cols[l];
            for (int b = 0; b < B; ++b) {
                docid[(size_t)l * B + b] = (int32_t)d[b];
                labels[(size_t)b * L + l] = y[b];
            }
        }
#pragma omp single nowait
        for (int k = 0; k < F; ++k) f32[nf + k] = 0.f;      // the PAD row (base_algorithm.py:148-149)
    }
    return 0;
}

// Pieces of ub200_pack_feed_host for a pipelined pack: the caller converts the feature rows in a few chunks and
// issues the H2D copy of chunk i while chunk i+1 is being converted.
extern "C" UB200_API int ub200_pack_ids_host(const float* const* docid_cols, const float* const* label_cols, int L,
                                             int B, void* dst, size_t dst_bytes) {
    UB_CHECK(dst && docid_cols && label_cols && L > 0 && B > 0, 2, "pack_ids_host: bad arguments");
    UB_CHECK(dst_bytes >= (size_t)8 * L * B, 3, "pack_ids_host: destination too small");
    int32_t* docid = reinterpret_cast<int32_t*>(dst);
    float* labels = reinterpret_cast<float*>(static_cast<char*>(dst) + (size_t)4 * L * B);
    for (int l = 0; l < L; ++l) {
        const float* d = docid_cols[l];
        const float* y = label_cols[l];
        for (int b = 0; b < B; ++b) {
            docid[(size_t)l * B + b] = (int32_t)d[b];
            labels[(size_t)b * L + l] = y[b];
        }
    }
    return 0;
}

extern "C" UB200_API int ub200_convert_f64_f32_host(const double* src, float* dst, size_t n, int n_threads) {
    UB_CHECK((src && dst) || n == 0, 2, "convert_f64_f32_host: null pointer");
    if (n_threads < 1) n_threads = 1;
    const long long nn = (long long)n, blk = 16384, nblk = (nn + blk - 1) / blk;
#pragma omp parallel for num_threads(n_threads) schedule(static)
    for (long long b = 0; b < nblk; ++b) {
        const long long lo = b * blk, hi = lo + blk < nn ? lo + blk : nn;
        for (long long i = lo; i < hi; ++i) dst[i] = (float)src[i];
    }
    return 0;
}
