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
#include <stdlib.h>
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

constexpr long long kBlock = 4096;       // elements per conversion block (32 KB in, 16 KB out): a block is ~2 us of one
                                         // core, so the short first group of a staged feed is done a few us after the start
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

// the same with ordinary stores: the staging buffers are reused every other step and small enough (5.6 MB at config 2)
// to stay in the last-level cache, where the PCIe read of the H2D copy finds them - no DRAM round trip for the fp32 copy
__attribute__((target("avx2"))) void cvt_avx2_cached(const double* s, float* d, long long n) {
    long long i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m128 a = _mm256_cvtpd_ps(_mm256_loadu_pd(s + i));
        const __m128 b = _mm256_cvtpd_ps(_mm256_loadu_pd(s + i + 4));
        const __m128 c = _mm256_cvtpd_ps(_mm256_loadu_pd(s + i + 8));
        const __m128 e = _mm256_cvtpd_ps(_mm256_loadu_pd(s + i + 12));
        _mm256_storeu_ps(d + i, _mm256_set_m128(b, a));
        _mm256_storeu_ps(d + i + 8, _mm256_set_m128(e, c));
    }
    for (; i < n; ++i) d[i] = (float)s[i];
}

typedef void (*cvt_fn)(const double*, float*, long long);
cvt_fn pick_cvt() {
    __builtin_cpu_init();
    if (!__builtin_cpu_supports("avx2")) return cvt_scalar;
    // streaming stores by default: measured on the B200 hosts, H2D copies out of a staging buffer that the CPUs left in
    // their caches (UB200_PACK_NT=0) run ~40 % slower than out of DRAM
    const char* nt = getenv("UB200_PACK_NT");
    return (nt && nt[0] == '0') ? cvt_avx2_cached : cvt_avx2;
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
    long long group_first[kMaxGroups + 1] = {0};   // group k = blocks [group_first[k], group_first[k + 1])
    bool caller_polls_only = false;
    int group_of(long long b) const {
        int k = 0;
        while (k + 1 < n_groups && b >= group_first[k + 1]) ++k;
        return k;
    }
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
        run(job, [](Slot&) {}, on_poll);
    }
    // `pre` runs on the caller right after the job has been published (the workers are already converting)
    template <typename P, typename F>
    void run(Job job, P&& pre, F&& on_poll) {
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
        pre(s);
        // with enough workers the caller only issues the copies (a cudaMemcpyAsync call costs ~4 us, during which a
        // converting caller would sit on a block the copy of its group waits for)
        if (!(job.caller_polls_only && job.n_workers > 0)) work(s, job, [&] { on_poll(s); });
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
        return s.group_done[k].load(std::memory_order_acquire) >= j.group_first[k + 1] - j.group_first[k];
    }

private:
    template <typename P>
    static void work(Slot& s, const Job& j, P&& after_block) {
        for (;;) {
            const long long b = s.next.fetch_add(1);
            if (b >= j.nblocks) break;
            const long long lo = b * kBlock, hi = lo + kBlock < j.n ? lo + kBlock : j.n;
            g_cvt(j.src + lo, j.dst + lo, hi - lo);
            s.group_done[j.group_of(b)].fetch_add(1, std::memory_order_release);
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
    if (j.nblocks < 16) w = 0;                                 // tiny inputs: not worth waking anybody
    if (w > 0) pool()->ensure(w);
    j.n_workers = w;
    if (n_groups > kMaxGroups) n_groups = kMaxGroups;
    if (n_groups > j.nblocks) n_groups = (int)j.nblocks;
    if (n_groups < 1) n_groups = 1;
    // Groups that double in size: the copy of the (small) first group starts almost at once and, because converting a
    // group is faster than copying the one before it (PCIe 5: ~107 us for 5.6 MB, conversion ~60 us on 15 threads), the
    // copy engine never waits again; every copy costs ~4 us of set-up, so few groups - 1/16, 2/16, 4/16, 9/16 by default.
    if (j.nblocks == 0) {
        j.n_groups = 0;
        return j;
    }
    long long unit = j.nblocks >> n_groups;
    if (unit < 1) unit = 1;
    // the first group is at most one block per worker (one ~2.5 us round): its copy leaves as early as possible
    long long first = (w > 0 && unit > w) ? w : unit;
    int k = 0;
    long long b = 0;
    j.group_first[0] = 0;
    while (k < n_groups - 1 && b + (k == 0 ? first : unit << k) < j.nblocks) {
        b += k == 0 ? first : unit << k;
        j.group_first[++k] = b;
    }
    j.group_first[++k] = j.nblocks;
    j.n_groups = k;
    return j;
}

// returns the number of ids outside [0, max_id] (max_id = the PAD row): the kernels gather feats[id] unchecked, where
// the reference's np.take raises IndexError (base_algorithm.py:150)
// slice `part` of `parts` of the ids / labels packing (positions [l0, l1) of the ids, queries [b0, b1) of the labels)
long long pack_ids_part(const float* const* docid_cols, const float* const* label_cols, int L, int B, void* dst,
                        long long max_id, int part, int parts) {
    int32_t* docid = reinterpret_cast<int32_t*>(dst);                                              // [L, B]
    float* labels = reinterpret_cast<float*>(static_cast<char*>(dst) + (size_t)4 * L * B);         // [B, L]
    long long bad = 0;
    const float hi = (float)max_id;
    const int l_begin = (int)((long long)L * part / parts), l_end = (int)((long long)L * (part + 1) / parts);
    const int q_begin = (int)((long long)B * part / parts), q_end = (int)((long long)B * (part + 1) / parts);
    for (int l = l_begin; l < l_end; ++l) {              // contiguous on both sides: vectorises
        const float* d = docid_cols[l];
        int32_t* o = docid + (size_t)l * B;
        long long bad_l = 0;
        for (int b = 0; b < B; ++b) {
            const float v = d[b];
            bad_l += !(v >= 0.f && v <= hi);
            o[b] = (int32_t)v;
        }
        bad += bad_l;
    }
    // the transpose [L][B] -> [B][L] in tiles of 32 queries: L read streams of 128 bytes each, contiguous writes
    for (int b0 = q_begin; b0 < q_end; b0 += 32) {
        const int b1 = b0 + 32 < q_end ? b0 + 32 : q_end;
        for (int b = b0; b < b1; ++b) {
            float* o = labels + (size_t)b * L;
            for (int l = 0; l < L; ++l) o[l] = label_cols[l][b];
        }
    }
    return bad;
}

long long pack_ids(const float* const* docid_cols, const float* const* label_cols, int L, int B, void* dst,
                   long long max_id) {
    return pack_ids_part(docid_cols, label_cols, L, B, dst, max_id, 0, 1);
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
// wall-clock stamps of the last stage_feed call (ns since its entry): [0] ids packed, [1] conversion job published,
// [2 + k] copy of group k issued, [30] job complete, [31] return; read with ub200_stage_timeline (diagnostics)
static long long g_stage_tl[32];
static inline long long now_ns() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
extern "C" UB200_API int ub200_stage_timeline(long long* out32) {
    for (int i = 0; i < 32; ++i) out32[i] = g_stage_tl[i];
    return 0;
}

static int stage_feed_impl(const double* feats, int n_docs, int F, const float* const* docid_cols,
                           const float* const* label_cols, int L, int B, void* pinned, size_t pinned_bytes, void* device,
                           int n_threads, int n_groups, cudaStream_t st) {
    HP_CHECK(pinned && device && docid_cols && label_cols && (feats || n_docs == 0), 2, "stage_feed: null pointer");
    HP_CHECK(L > 0 && B > 0 && F > 0 && n_docs >= 0, 1, "stage_feed: bad sizes");
    const size_t need = ub200_feed_bytes(n_docs, F, L, B);
    HP_CHECK(pinned_bytes >= need, 3, "stage_feed: staging buffer too small (%zu < %zu)", pinned_bytes, need);
    char* hb = static_cast<char*>(pinned);
    char* db = static_cast<char*>(device);
    const size_t off_f = align_up((size_t)8 * L * B, 256);
    float* f32 = reinterpret_cast<float*>(hb + off_f);
    const long long nf = (long long)n_docs * F;
    cudaError_t err = cudaSuccess;
    auto copy = [&](size_t b0, size_t b1) {
        if (err == cudaSuccess && b1 > b0) err = cudaMemcpyAsync(db + b0, hb + b0, b1 - b0, cudaMemcpyHostToDevice, st);
    };
    const long long t_in = now_ns();
    for (int i = 0; i < 32; ++i) g_stage_tl[i] = 0;
    long long bad = 0;
    if (nf == 0) {
        bad = pack_ids(docid_cols, label_cols, L, B, pinned, n_docs);
        memset(f32, 0, sizeof(float) * (size_t)F);           // the PAD row
        copy(0, need);
    } else {
        // The workers start converting at once.  The caller packs the ids / labels block in eight slices, polling for
        // finished groups in between (the first group of feature rows is usually on its way before the ids are
        // packed), copies it, and then converts / polls with the others.
        Job job = make_job(feats, f32, nf, n_threads, n_groups);
        // with a full set of workers the caller only packs the ids and issues the copies (measured: the first copy
        // leaves ~12 us earlier than when the caller converts blocks between polls); UB200_PACK_ISSUER=0 / 1 forces it
        static const int issuer = [] { const char* v = getenv("UB200_PACK_ISSUER"); return v ? (v[0] == '1' ? 1 : 0) : -1; }();
        job.caller_polls_only = issuer == 1 || (issuer == -1 && job.n_workers >= 7);
        int issued = 0;
        auto poll = [&](Slot& s) {
            while (issued < job.n_groups && Pool::group_complete(s, job, issued)) {
                const long long e0 = job.group_first[issued] * kBlock;
                long long e1 = job.group_first[issued + 1] * kBlock;
                const bool last = issued == job.n_groups - 1;
                if (e1 > nf) e1 = nf;
                copy(off_f + 4 * (size_t)e0, last ? need : off_f + 4 * (size_t)e1);
                if (issued < 28) g_stage_tl[2 + issued] = now_ns() - t_in;
                ++issued;
            }
        };
        auto pack_head = [&](Slot& s) {
            memset(f32 + nf, 0, sizeof(float) * (size_t)F);      // the PAD row (travels with the last group)
            for (int part = 0; part < 8; ++part) {
                bad += pack_ids_part(docid_cols, label_cols, L, B, pinned, n_docs, part, 8);
                poll(s);
            }
            copy(0, (size_t)8 * L * B);
            g_stage_tl[0] = now_ns() - t_in;
        };
        pool()->run(job, pack_head, poll);
    }
    g_stage_tl[31] = now_ns() - t_in;
    HP_CHECK(bad == 0, 5, "stage_feed: %lld document ids outside [0, %d]", bad, n_docs);
    HP_CHECK(err == cudaSuccess, 100, "stage_feed: cudaMemcpyAsync failed: %s", cudaGetErrorString(err));
    return 0;
}

extern "C" UB200_API int ub200_stage_feed(const double* feats, int n_docs, int F, const float* const* docid_cols,
                                          const float* const* label_cols, int L, int B, void* pinned,
                                          size_t pinned_bytes, void* device, int n_threads, int n_groups,
                                          void* stream) {
    return stage_feed_impl(feats, n_docs, F, docid_cols, label_cols, L, B, pinned, pinned_bytes, device, n_threads,
                           n_groups, static_cast<cudaStream_t>(stream));
}

// ---- double-buffered staging: the copies of step i run on their own stream beside the kernels of step i - 1 ----------
// The caller alternates between two (pinned, device) buffer pairs.  `slot_free` (may be null) is an event recorded on the
// compute stream behind the last kernel that reads this device buffer: the copy stream waits for it before overwriting.
// After the last copy `ready` is recorded on the copy stream and the compute stream is made to wait for it, so kernels
// launched on the compute stream afterwards see the complete feed - without the copies queueing behind the backward
// pass and optimizer step of the previous batch that are still running there.
static int pipeline_begin(cudaStream_t copy, void* slot_free) {
    if (slot_free) {
        cudaError_t e = cudaStreamWaitEvent(copy, static_cast<cudaEvent_t>(slot_free), 0);
        HP_CHECK(e == cudaSuccess, 100, "staging: cudaStreamWaitEvent(copy, slot_free) failed: %s", cudaGetErrorString(e));
    }
    return 0;
}
static int pipeline_end(cudaStream_t copy, cudaStream_t compute, void* ready) {
    HP_CHECK(ready != nullptr, 2, "staging: null `ready` event");
    cudaError_t e = cudaEventRecord(static_cast<cudaEvent_t>(ready), copy);
    HP_CHECK(e == cudaSuccess, 100, "staging: cudaEventRecord failed: %s", cudaGetErrorString(e));
    e = cudaStreamWaitEvent(compute, static_cast<cudaEvent_t>(ready), 0);
    HP_CHECK(e == cudaSuccess, 100, "staging: cudaStreamWaitEvent(compute, ready) failed: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" UB200_API int ub200_stage_feed_pipelined(const double* feats, int n_docs, int F,
                                                    const float* const* docid_cols, const float* const* label_cols,
                                                    int L, int B, void* pinned, size_t pinned_bytes, void* device,
                                                    int n_threads, int n_groups, void* copy_stream,
                                                    void* compute_stream, void* slot_free, void* ready) {
    cudaStream_t copy = static_cast<cudaStream_t>(copy_stream);
    if (int rc = pipeline_begin(copy, slot_free)) return rc;
    if (int rc = stage_feed_impl(feats, n_docs, F, docid_cols, label_cols, L, B, pinned, pinned_bytes, device,
                                 n_threads, n_groups, copy))
        return rc;
    return pipeline_end(copy, static_cast<cudaStream_t>(compute_stream), ready);
}

// the same for a device-resident data set: only the ids / labels block (ub200_pack_ids_host layout) crosses PCIe
extern "C" UB200_API int ub200_stage_ids_pipelined(const float* const* docid_cols, const float* const* label_cols,
                                                   int L, int B, int max_id, void* pinned, size_t pinned_bytes,
                                                   void* device, void* copy_stream, void* compute_stream,
                                                   void* slot_free, void* ready) {
    HP_CHECK(pinned && device && docid_cols && label_cols, 2, "stage_ids: null pointer");
    HP_CHECK(L > 0 && B > 0, 1, "stage_ids: bad sizes");
    const size_t need = (size_t)8 * L * B;
    HP_CHECK(pinned_bytes >= need, 3, "stage_ids: staging buffer too small (%zu < %zu)", pinned_bytes, need);
    const long long bad = pack_ids(docid_cols, label_cols, L, B, pinned, max_id);
    HP_CHECK(bad == 0, 5, "stage_ids: %lld document ids outside [0, %d]", bad, max_id);
    cudaStream_t copy = static_cast<cudaStream_t>(copy_stream);
    if (int rc = pipeline_begin(copy, slot_free)) return rc;
    cudaError_t e = cudaMemcpyAsync(device, pinned, need, cudaMemcpyHostToDevice, copy);
    HP_CHECK(e == cudaSuccess, 100, "stage_ids: cudaMemcpyAsync failed: %s", cudaGetErrorString(e));
    return pipeline_end(copy, static_cast<cudaStream_t>(compute_stream), ready);
}

// records `event` on `stream` (the engine marks "every kernel launched so far has been enqueued" with it)
extern "C" UB200_API int ub200_event_record(void* event, void* stream) {
    HP_CHECK(event != nullptr, 2, "event_record: null event");
    cudaError_t e = cudaEventRecord(static_cast<cudaEvent_t>(event), static_cast<cudaStream_t>(stream));
    HP_CHECK(e == cudaSuccess, 100, "event_record: %s", cudaGetErrorString(e));
    return 0;
}
