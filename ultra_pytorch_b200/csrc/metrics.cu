// N2 - validation metrics on the device: per-list NDCG@n / ERR@n / MRR from the ranker's scores where it left them.
//
// Replaces (reference): remove_padding_for_metric_eval ultra/learning_algorithm/base_algorithm.py:88-116 (a Python loop
// over the L positions) and, per list, the sort / gather / cumsum chains of ultra/utils/metrics.py:191-221 (DCG),
// :224-265 (label validation), :268-298 (MRR), :300-336 (ERR), :456-495 (NDCG).  Only the batch means stay on the host
// (ultra_pytorch_b200/metrics.py), on B x n values instead of the B x L scores.
//
// One CTA per ranked list.  The arithmetic follows the reference's torch-CPU semantics step by step so that NDCG and
// MRR are bit-identical given identical scores:
//   - scores of PAD documents (docid == n_docs) become -100000; labels < 0 become 0 and their score min - 1e-6;
//   - descending STABLE order by counting (rank_i = #{p_j > p_i} + #{j < i : p_j == p_i});
//   - gain = 2^y - 1 (exact for the integer grades the data sets use; other labels raise a flag and the caller falls
//     back), discount = 1 / log2(r + 2) taken from a table the HOST computed with the reference's own torch expression;
//   - torch.cumsum / torch.cumprod on CPU accumulate float data in DOUBLE and round every prefix to float: so do we;
//   - DCG / ideal DCG -> safe_div; ERR = sum_r rel_r * nonrel_r * (1/r) [r <= n] added in rank order (torch.sum's own
//     vectorised order is not reproducible, ERR agrees to ~1e-7); MRR = max_r [y_r >= 1] / r.
#include "common.cuh"

namespace ub200 {

constexpr int kMetThreads = 128;
constexpr int kMaxTopn = 8;

struct MetArgs {
    const float* scores;        // [B, L]
    const float* labels;        // [B, L]
    const int32_t* docid;       // [L, B] or nullptr (no PAD masking)
    const float* discount;      // [L] 1 / log2(r + 2), host-computed
    int n_docs, B, L;
    int n_topn;
    int topn[kMaxTopn];         // already clipped to L
    float max_label_pow;        // 2^MAX_LABEL
    float* out;                 // [B, 2 * n_topn + 1]: ndcg@n.. | err@n.. | mrr
    int* flag;                  // set to 1 when a label is not an integer in [0, 30]
};

__global__ void __launch_bounds__(kMetThreads) rank_metrics_kernel(MetArgs a) {
    griddep_launch();
    griddep_wait();
    extern __shared__ float sm[];
    const int L = a.L, b = blockIdx.x, tid = threadIdx.x;
    float* p = sm;                 // predictions after masking
    float* y = p + L;              // labels after validation
    float* ys = y + L;             // labels in predicted order
    float* yi = ys + L;            // labels in ideal order
    __shared__ float s_min[kMetThreads / 32];
    __shared__ int s_bad;
    if (tid == 0) s_bad = 0;
    float mn = INFINITY;
    for (int l = tid; l < L; l += kMetThreads) {
        float s = a.scores[(size_t)b * L + l];
        if (a.docid && a.docid[(size_t)l * a.B + b] == a.n_docs) s = -100000.0f;
        p[l] = s;
        mn = fminf(mn, s);
    }
    mn = -warp_max(-mn);
    if ((tid & 31) == 0) s_min[tid >> 5] = mn;
    __syncthreads();
    mn = fminf(fminf(s_min[0], s_min[1]), fminf(s_min[2], s_min[3]));
    const float invalid_score = -1e-6f + mn;
    bool bad = false;
    for (int l = tid; l < L; l += kMetThreads) {
        float v = a.labels[(size_t)b * L + l];
        if (!(v >= 0.f)) {
            v = 0.f;
            p[l] = invalid_score;
        }
        bad = bad || v != floorf(v) || v > 30.f;
        y[l] = v;
    }
    if (bad) s_bad = 1;
    __syncthreads();
    for (int i = tid; i < L; i += kMetThreads) {
        const float pi = p[i], vi = y[i];
        int rp = 0, ri = 0;
        for (int j = 0; j < L; ++j) {
            const float pj = p[j], vj = y[j];
            rp += (pj > pi) || (pj == pi && j < i);
            ri += (vj > vi) || (vj == vi && j < i);
        }
        ys[rp] = vi;
        yi[ri] = vi;
    }
    __syncthreads();
    float* out = a.out + (size_t)b * (2 * a.n_topn + 1);
    if (tid == 0) {
        if (s_bad) *a.flag = 1;
        int max_n = 0;
        for (int q = 0; q < a.n_topn; ++q) max_n = max(max_n, a.topn[q]);
        // DCG and ideal DCG: prefix sums accumulated in double, every prefix rounded to float (torch CPU cumsum)
        double cd = 0.0, ci = 0.0;
        int q = 0;
        float dcg[kMaxTopn], idcg[kMaxTopn];
        for (int r = 0; r < max_n; ++r) {
            const float disc = a.discount[r];
            const float gd = (ldexpf(1.0f, (int)ys[r]) - 1.0f) * disc;
            const float gi = (ldexpf(1.0f, (int)yi[r]) - 1.0f) * disc;
            cd += (double)gd;
            ci += (double)gi;
            for (q = 0; q < a.n_topn; ++q)
                if (a.topn[q] - 1 == r) {
                    dcg[q] = (float)cd;
                    idcg[q] = (float)ci;
                }
        }
        for (q = 0; q < a.n_topn; ++q) out[q] = idcg[q] == 0.f ? 0.f : dcg[q] / idcg[q];
    } else if (tid == 32) {
        // ERR: relevance, running product of (1 - relevance) in double rounded per step (torch CPU cumprod)
        float err[kMaxTopn];
        for (int q = 0; q < a.n_topn; ++q) err[q] = 0.f;
        double cp = 1.0;
        for (int r = 0; r < L; ++r) {
            const float rel = (ldexpf(1.0f, (int)ys[r]) - 1.0f) / a.max_label_pow;
            const float om = 1.0f - rel;
            cp *= (double)om;
            const float nonrel = (float)cp / om;
            const float rr = 1.0f / (float)(r + 1);
            const float base = rel * nonrel;
            for (int q = 0; q < a.n_topn; ++q) {
                const float rrq = (r < a.topn[q]) ? rr : rr * 0.0f;
                err[q] += base * rrq * 1.0f;
            }
        }
        for (int q = 0; q < a.n_topn; ++q) out[a.n_topn + q] = err[q];
    } else if (tid == 64) {
        float m = 0.f;      // every candidate is >= 0 and the reference's max runs over the whole list
        for (int r = 0; r < L; ++r) m = fmaxf(m, (ys[r] >= 1.0f ? 1.0f : 0.0f) * (1.0f / (float)(r + 1)));
        out[2 * a.n_topn] = m;
    }
}

}  // namespace ub200

using namespace ub200;

extern "C" UB200_API int ub200_rank_metrics(const float* scores, const float* labels, const int32_t* docid, int n_docs,
                                            int B, int L, const float* discount, const int* topn, int n_topn,
                                            float max_label, float* out, int* flag, void* stream) {
    UB_CHECK(scores && labels && discount && topn && out && flag, 2, "rank_metrics: null pointer");
    UB_CHECK(B > 0 && L > 0 && n_topn > 0 && n_topn <= kMaxTopn, 1, "rank_metrics: bad sizes B=%d L=%d n_topn=%d", B, L,
             n_topn);
    const size_t smem = sizeof(float) * 4 * (size_t)L;
    UB_CHECK(smem <= 200 * 1024, 4, "rank_metrics: list length %d too large", L);
    MetArgs a;
    a.scores = scores; a.labels = labels; a.docid = docid; a.discount = discount;
    a.n_docs = n_docs; a.B = B; a.L = L; a.n_topn = n_topn;
    for (int q = 0; q < n_topn; ++q) a.topn[q] = topn[q] < L ? topn[q] : L;
    a.max_label_pow = exp2f(max_label);
    a.out = out; a.flag = flag;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(rank_metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(rank_metrics_kernel, B, kMetThreads, smem, static_cast<cudaStream_t>(stream), a);
    UB_LAUNCH_CHECK("rank_metrics_kernel");
    return 0;
}
