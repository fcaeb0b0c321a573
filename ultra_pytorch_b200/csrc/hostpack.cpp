// Host side of the boundary: pack one input_feed into the pinned staging buffer and ship it to the device.
// Replaces the numpy work of BaseAlgorithm.create_input_feed / get_ranking_scores
// (base_algorithm.py:148-152 concat + np.take, :176-186 label transpose / docid conversion, DNN.py:72-73 f64 -> f32).
//
// The f64 -> f32 conversion of the feature rows (11 MB in, 5.6 MB out per 256-query batch at config 2) is the
// dominant host cost of a training step, so it runs on a persistent pool of spinning worker threads (no fork/join
// wake-up per call), with AVX2 conversion and non-temporal stores into the pinned buffer (no read-for-ownership, no
// cache pollution: the DMA engine is the only reader), and the H2D copy of every finished group of blocks is issued
// while the remaining blocks are still being converted.
//
// Plain C++ (g++): no device code in this file.
#include <cuda_runtime_api.h>
#include <immintrin.h>
#include <sched.h>
#include <stdint.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/ultra_b200.h"

namespace ub200 {
void set_error(const char* fmt, ...);   // optim.cu
}

namespace {

#define HP_CHECK(cond, code, ...)            \
    do {                                     \
        if (!(cond)) {                       \
            ub200::set_error(__VA_ARGS__);   \
            return (code);                   \
        }                                    \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr long long kBlock = 16384;      // elements per conversion block (128 KB in, 64 KB out)
constexpr int kMaxGroups = 16;

// ---- f64 -> f32 ---------------------------------------------------------------------------------------------
void cvt_scalar(const double* s, float* d, long long n) {
    for (long long i = 0; i < n; ++i) d[i] = (float)s[i];
}

__attribute__((target("avx2"))) void cvt_avx2(const double* s, float* d, long long n) {
    long long i = 0;
    // head: scalar until d is 32-byte aligned (streaming stores need it)
    while (i < n && (reinterpret_cast<uintptr_t>(d + i) & 31)) {
        d[i] = (float)s[i];
        ++i;
    }
    for (; i + 16 <= n; i += 16) {
        const __m128 a = _mm256_cvtpd_ps(_mm256_loadu_pd(s + i));           // round-to-nearest-even == (float)x
        const __m128 b = _mm256_cvtpd_ps(_mm256_loadu_pd(s + i + 4));
        const __m128 c = _mm256_cvtpd_ps(_mm256_loadu_pd(s + i + 8));
        const __m128 e = _mm256_cvtpd_ps(_mm256_loadu_pd(s + i + 12));
        _mm256_stream_ps(d + i, _mm256_set_m128(b, a));
        _mm256_stream_ps(d + i + 8, _mm256_set_m128(e, c));
    }
    for (; i < n; ++i) d[i] = (float)s[i];
    _mm_sfence();
}

typedef void (*cvt_fn)(const double*, float*, long long);
cvt_fn pick_cvt() {
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx2") ? cvt_avx2 : cvt_scalar;
}
const cvt_fn g_cvt = pick_cvt();

// ---- persistent worker pool --------------------------------------------------------------------------------------
// Jobs are double-buffered: job g lives in slot g & 1, so publishing job g+1 never touches what the workers of job g
// read; a slot is rewritten only after every worker that entered it has left (active == 0), and a worker reads a
// slot only if the generation it saw is still current after it registered itself (entry re-check).
struct Job {
    const double* src = nullptr;
    float* dst = nullptr;
    long long n = 0, nblocks = 0;
    int n_workers = 0;                      // workers (besides the caller) that should join
    int n_groups = 1;
    long long blocks_per_group = 1;
};
struct Slot {
    Job job;
    std::atomic<long long> next{0};
    std::atomic<int> active{0};
    std::atomic<long long> group_done[kMaxGroups];
};

class Pool {
public:
    explicit Pool(int max_workers) : max_(max_workers) {
        for (int s = 0; s < 2; ++s)
            for (int g = 0; g < kMaxGroups; ++g) slots_[s].group_done[g].store(0);
    }
    int max_workers() const { return max_; }
    // workers are created on demand: a process that packs with 2 threads (8 ranks sharing 16 cores) must not keep 15
    // spinning threads around
    void ensure(int n) {
        if (n > max_) n = max_;
        std::lock_guard<std::mutex> lk(call_mu_);
        while (n_ < n) {
            std::thread(&Pool::worker, this, n_).detach();
            ++n_;
        }
    }

    // runs `job` on the caller + up to job.n_workers pool threads; `on_poll` is called by the caller after each
    // block it converts and while it waits (used to issue the H2D copies of finished groups)
    template <typename F>
    void run(Job job, F&& on_poll) {
        std::lock_guard<std::mutex> call_lock(call_mu_);        // one job at a time
        const uint64_t g = gen_.load() + 1;
        Slot& s = slots_[g & 1];
        while (s.active.load() != 0) _mm_pause();               // stragglers of job g-2
        s.job = job;
        s.next.store(0);
        for (int k = 0; k < kMaxGroups; ++k) s.group_done[k].store(0);
        gen_.store(g);
        if (job.n_workers > 0 && sleepers_.load() > 0) {
            std::lock_guard<std::mutex> lk(mu_);
            cv_.notify_all();
        }
        work(s, job, [&] { on_poll(s); });
        for (;;) {
            bool all = true;
            for (int k = 0; k < job.n_groups; ++k) all = all && group_complete(s, job, k);
            on_poll(s);
            if (all) break;
            _mm_pause();
            sched_yield();
        }
    }
    static bool group_complete(const Slot& s, const Job& j, int k) {
        const long long b0 = k * j.blocks_per_group;
        long long b1 = b0 + j.blocks_per_group;
        if (b1 > j.nblocks) b1 = j.nblocks;
        return s.group_done[k].load(std::memory_order_acquire) >= b1 - b0;
    }

private:
    template <typename P>
    static void work(Slot& s, const Job& j, P&& after_block) {
        for (;;) {
            const long long b = s.next.fetch_add(1);
            if (b >= j.nblocks) break;
            const long long lo = b * kBlock, hi = lo + kBlock < j.n ? lo + kBlock : j.n;
            g_cvt(j.src + lo, j.dst + lo, hi - lo);
            s.group_done[b / j.blocks_per_group].fetch_add(1, std::memory_order_release);
            after_block();
        }
    }
    void worker(int index) {
        uint64_t seen = 0;
        for (;;) {
            // ---- wait for a new generation: spin for ~2 ms, then sleep ----
            uint64_t g;
            unsigned spins = 0;
            auto t0 = std::chrono::steady_clock::now();
            while ((g = gen_.load()) == seen) {
                _mm_pause();
                if ((++spins & 63) == 0) sched_yield();      // oversubscribed cores go to whoever has work
                if ((spins & 1023) == 0 &&
                    std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(2000)) {
                    std::unique_lock<std::mutex> lk(mu_);
                    sleepers_.fetch_add(1);
                    cv_.wait(lk, [&] { return gen_.load() != seen; });
                    sleepers_.fetch_sub(1);
                    t0 = std::chrono::steady_clock::now();
                }
            }
            Slot& s = slots_[g & 1];
            s.active.fetch_add(1);
            if (gen_.load() != g) {          // the job changed while we were entering: do not read the slot
                s.active.fetch_sub(1);
                continue;
            }
            seen = g;
            const Job j = s.job;
            if (index < j.n_workers) work(s, j, [] {});
            s.active.fetch_sub(1);
        }
    }

    const int max_;
    int n_ = 0;
    Slot slots_[2];
    std::atomic<uint64_t> gen_{0};
    std::atomic<int> sleepers_{0};
    std::mutex mu_, call_mu_;
    std::condition_variable cv_;
};

Pool* pool() {
    // never destroyed (the workers are detached); re-created in a forked child, whose threads did not survive
    static Pool* p = nullptr;
    static pid_t owner = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (p == nullptr || owner != getpid()) {
        unsigned hc = std::thread::hardware_concurrency();
        int n = hc > 1 ? (int)hc - 1 : 0;
        if (n > 31) n = 31;
        p = new Pool(n);                      // upper bound; threads are spawned by ensure()
        owner = getpid();
    }
    return p;
}

Job make_job(const double* src, float* dst, long long n, int n_threads, int n_groups) {
    Job j;
    j.src = src;
    j.dst = dst;
    j.n = n;
    j.nblocks = (n + kBlock - 1) / kBlock;
    int w = (n_threads < 1 ? 1 : n_threads) - 1;
    if (w > pool()->max_workers()) w = pool()->max_workers();
    if (j.nblocks < 4) w = 0;                                  // tiny inputs: not worth waking anybody
    if (w > 0) pool()->ensure(w);
    j.n_workers = w;
    if (n_groups > kMaxGroups) n_groups = kMaxGroups;
    if (n_groups > j.nblocks) n_groups = (int)j.nblocks;
    if (n_groups < 1) n_groups = 1;
    j.n_groups = n_groups;
    j.blocks_per_group = (j.nblocks + n_groups - 1) / n_groups;
    if (j.blocks_per_group < 1) j.blocks_per_group = 1;
    j.n_groups = j.nblocks ? (int)((j.nblocks + j.blocks_per_group - 1) / j.blocks_per_group) : 0;
    return j;
}

// returns the number of ids outside [0, max_id] (max_id = the PAD row): the kernels gather feats[id] unchecked, where
// the reference's np.take raises IndexError (base_algorithm.py:150)
long long pack_ids(const float* const* docid_cols, const float* const* label_cols, int L, int B, void* dst,
                   long long max_id) {
    int32_t* docid = reinterpret_cast<int32_t*>(dst);                                              // [L, B]
    float* labels = reinterpret_cast<float*>(static_cast<char*>(dst) + (size_t)4 * L * B);         // [B, L]
    long long bad = 0;
    for (int l = 0; l < L; ++l) {
        const float* d = docid_cols[l];
        const float* y = label_cols[l];
        for (int b = 0; b < B; ++b) {
            const float v = d[b];
            bad += !(v >= 0.f && v <= (float)max_id);
            docid[(size_t)l * B + b] = (int32_t)v;
            labels[(size_t)b * L + l] = y[b];
        }
    }
    return bad;
}

}  // namespace

extern "C" UB200_API size_t ub200_feed_bytes(int n_docs, int F, int L, int B) {
    const size_t off_f = align_up((size_t)8 * L * B, 256);
    return off_f + sizeof(float) * (size_t)(n_docs + 1) * F;
}

extern "C" UB200_API int ub200_convert_f64_f32_host(const double* src, float* dst, size_t n, int n_threads) {
    HP_CHECK((src && dst) || n == 0, 2, "convert_f64_f32_host: null pointer");
    if (n == 0) return 0;
    pool()->run(make_job(src, dst, (long long)n, n_threads, 1), [](Slot&) {});
    return 0;
}

extern "C" UB200_API int ub200_pack_ids_host(const float* const* docid_cols, const float* const* label_cols, int L,
                                             int B, int max_id, void* dst, size_t dst_bytes) {
    HP_CHECK(dst && docid_cols && label_cols && L > 0 && B > 0, 2, "pack_ids_host: bad arguments");
    HP_CHECK(dst_bytes >= (size_t)8 * L * B, 3, "pack_ids_host: destination too small");
    const long long bad = pack_ids(docid_cols, label_cols, L, B, dst, max_id);
    HP_CHECK(bad == 0, 5, "pack_ids_host: %lld document ids outside [0, %d] (the feed indexes rows the feature matrix does not have)", bad, max_id);
    return 0;
}

extern "C" UB200_API int ub200_pack_feed_host(const double* feats, int n_docs, int F, const float* const* docid_cols,
                                              const float* const* label_cols, int L, int B, void* dst,
                                              size_t dst_bytes, int n_threads) {
    HP_CHECK(dst && docid_cols && label_cols && (feats || n_docs == 0), 2, "pack_feed_host: null pointer");
    HP_CHECK(L > 0 && B > 0 && F > 0 && n_docs >= 0, 1, "pack_feed_host: bad sizes");
    const size_t need = ub200_feed_bytes(n_docs, F, L, B);
    HP_CHECK(dst_bytes >= need, 3, "pack_feed_host: destination too small (%zu < %zu)", dst_bytes, need);
    char* base = static_cast<char*>(dst);
    float* f32 = reinterpret_cast<float*>(base + align_up((size_t)8 * L * B, 256));   // [n_docs + 1, F]
    const long long bad = pack_ids(docid_cols, label_cols, L, B, dst, n_docs);
    HP_CHECK(bad == 0, 5, "pack_feed_host: %lld document ids outside [0, %d]", bad, n_docs);
    const long long nf = (long long)n_docs * F;
    memset(f32 + nf, 0, sizeof(float) * (size_t)F);      // the PAD row (base_algorithm.py:148-149)
    if (nf) pool()->run(make_job(feats, f32, nf, n_threads, 1), [](Slot&) {});
    return 0;
}

// pack (as ub200_pack_feed_host) into the PINNED buffer `pinned` and copy it to `device` on `stream`, pipelined: the
// H2D copy of each finished group of blocks is issued while the rest is still being converted.  Returns after the
// last copy has been ENQUEUED (stream order makes the data visible to the kernels launched after it); `pinned` may
// be rewritten once the stream has passed the last copy (the caller syncs once per step for the loss anyway).
extern "C" UB200_API int ub200_stage_feed(const double* feats, int n_docs, int F, const float* const* docid_cols,
                                          const float* const* label_cols, int L, int B, void* pinned,
                                          size_t pinned_bytes, void* device, int n_threads, int n_groups,
                                          void* stream) {
    HP_CHECK(pinned && device && docid_cols && label_cols && (feats || n_docs == 0), 2, "stage_feed: null pointer");
    HP_CHECK(L > 0 && B > 0 && F > 0 && n_docs >= 0, 1, "stage_feed: bad sizes");
    const size_t need = ub200_feed_bytes(n_docs, F, L, B);
    HP_CHECK(pinned_bytes >= need, 3, "stage_feed: staging buffer too small (%zu < %zu)", pinned_bytes, need);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* hb = static_cast<char*>(pinned);
    char* db = static_cast<char*>(device);
    const size_t off_f = align_up((size_t)8 * L * B, 256);
    float* f32 = reinterpret_cast<float*>(hb + off_f);
    const long long nf = (long long)n_docs * F;
    cudaError_t err = cudaSuccess;
    auto copy = [&](size_t b0, size_t b1) {
        if (err == cudaSuccess && b1 > b0) err = cudaMemcpyAsync(db + b0, hb + b0, b1 - b0, cudaMemcpyHostToDevice, st);
    };
    const long long bad = pack_ids(docid_cols, label_cols, L, B, pinned, n_docs);
    HP_CHECK(bad == 0, 5, "stage_feed: %lld document ids outside [0, %d]", bad, n_docs);
    memset(f32 + nf, 0, sizeof(float) * (size_t)F);      // the PAD row
    copy(0, (size_t)8 * L * B);
    if (nf == 0) {
        copy(off_f, need);
    } else {
        const Job job = make_job(feats, f32, nf, n_threads, n_groups);
        int issued = 0;
        pool()->run(job, [&](Slot& s) {
            while (issued < job.n_groups && Pool::group_complete(s, job, issued)) {
                const long long e0 = (long long)issued * job.blocks_per_group * kBlock;
                long long e1 = e0 + job.blocks_per_group * kBlock;
                const bool last = issued == job.n_groups - 1;
                if (e1 > nf) e1 = nf;
                copy(off_f + 4 * (size_t)e0, last ? need : off_f + 4 * (size_t)e1);
                ++issued;
            }
        });
    }
    HP_CHECK(err == cudaSuccess, 100, "stage_feed: cudaMemcpyAsync failed: %s", cudaGetErrorString(err));
    return 0;
}
