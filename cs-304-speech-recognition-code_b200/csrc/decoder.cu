// Host-buffer decode entry point: a self-contained C pipeline (no torch, no Python) that takes raw PCM in
// HOST memory and returns the decoded word ids, i.e. the whole of
//   MFCC.batch -> HiddenMarkovModelInference.predict  (mfcc.py:71-84, hidden_markov_model.py:458-461)
// for a batch of utterances.  The decoder object owns the device copies of the model tables, a
// workspace that grows on demand and two streams: the batch is cut into chunks of whole utterances and
// the host->device copy of chunk c+1 (copy stream) overlaps MFCC, emission scoring and Viterbi of
// chunk c (compute stream); only the word-id tables (and optionally paths / scores) travel back.
#include "common.cuh"
#include <vector>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <stdlib.h>
#include <string.h>
#include <immintrin.h>
#include <pthread.h>
#include <sched.h>


namespace loe {

// ------------------------------------------------------------------------------------------------
// Lossless narrowing of float32 PCM on the host.  The reference hands the decoder float32 copies of int16 WAV
// samples (ti_digits.py:85-139), so half of the PCIe bytes are zeros in disguise.  Worker threads convert a chunk
// to int16 in pinned staging while the previous chunk is on the wire and VERIFY every sample (convert back,
// compare): a chunk with a single sample that is not an int16 value travels as float32 like before.  The MFCC
// kernel produces bit-identical features from either format (tests: f32 vs s16), so results do not depend on it.
// ------------------------------------------------------------------------------------------------
static bool narrow_range_sse2(const float* src, int16_t* dst, int64_t n) {
    __m128i bad = _mm_setzero_si128();
    int64_t i = 0;
    for (; i + 8 <= n; i += 8) {
        const __m128 a = _mm_loadu_ps(src + i), b = _mm_loadu_ps(src + i + 4);
        const __m128i p = _mm_packs_epi32(_mm_cvttps_epi32(a), _mm_cvttps_epi32(b));       // saturating
        const __m128 ra = _mm_cvtepi32_ps(_mm_srai_epi32(_mm_unpacklo_epi16(p, p), 16));
        const __m128 rb = _mm_cvtepi32_ps(_mm_srai_epi32(_mm_unpackhi_epi16(p, p), 16));
        bad = _mm_or_si128(bad, _mm_castps_si128(_mm_or_ps(_mm_cmpneq_ps(ra, a), _mm_cmpneq_ps(rb, b))));   // NaN != anything
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), p);
    }
    bool ok = _mm_movemask_epi8(bad) == 0;
    for (; i < n; ++i) {
        const float v = src[i];
        const bool in = v >= -32768.0f && v <= 32767.0f;
        const int16_t t = in ? (int16_t)(int32_t)v : (int16_t)0;
        ok = ok && in && (float)t == v;
        dst[i] = t;
    }
    return ok;
}

__attribute__((target("avx2"))) static bool narrow_range_avx2(const float* src, int16_t* dst, int64_t n) {
    __m256i bad = _mm256_setzero_si256();
    int64_t i = 0;
    // the int16 copy is written once and read by the DMA engine only: streaming stores (no read-for-ownership of
    // the destination lines) when the destination is 32-byte aligned
    const bool stream = (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
    for (; i + 16 <= n; i += 16) {
        const __m256 a = _mm256_loadu_ps(src + i), b = _mm256_loadu_ps(src + i + 8);
        // packs works per 128-bit lane: [a0-3 b0-3 | a4-7 b4-7] -> permute the 64-bit quarters back into order
        const __m256i p = _mm256_permute4x64_epi64(_mm256_packs_epi32(_mm256_cvttps_epi32(a), _mm256_cvttps_epi32(b)), 0xD8);
        const __m256 ra = _mm256_cvtepi32_ps(_mm256_cvtepi16_epi32(_mm256_castsi256_si128(p)));
        const __m256 rb = _mm256_cvtepi32_ps(_mm256_cvtepi16_epi32(_mm256_extracti128_si256(p, 1)));
        bad = _mm256_or_si256(bad, _mm256_castps_si256(_mm256_or_ps(_mm256_cmp_ps(ra, a, _CMP_NEQ_UQ), _mm256_cmp_ps(rb, b, _CMP_NEQ_UQ))));
        if (stream) _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), p);
        else _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), p);
    }
    if (stream) _mm_sfence();
    const bool ok = _mm256_testz_si256(bad, bad) != 0;
    return narrow_range_sse2(src + i, dst + i, n - i) && ok;
}

static bool narrow_range(const float* src, int16_t* dst, int64_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    return avx2 ? narrow_range_avx2(src, dst, n) : narrow_range_sse2(src, dst, n);
}

// persistent worker threads: run(f) calls f(worker index) on every worker and returns when all are done
class Pool {
public:
    // cpus: the CPUs the workers may run on (empty: wherever the process may)
    Pool(int n, const std::vector<int>& cpus) : n_(n) {
        for (int i = 0; i < n; ++i) th_.emplace_back([this, i] { loop(i); });
        if (!cpus.empty()) {
            cpu_set_t set;
            CPU_ZERO(&set);
            for (int c : cpus) if (c >= 0 && c < CPU_SETSIZE) CPU_SET(c, &set);
            for (auto& t : th_) pthread_setaffinity_np(t.native_handle(), sizeof(set), &set);   // best effort
        }
    }
    ~Pool() {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    int size() const { return n_; }
    void run(const std::function<void(int)>& f) {
        std::unique_lock<std::mutex> l(m_);
        job_ = &f; pending_ = n_; ++gen_;
        cv_.notify_all();
        done_.wait(l, [this] { return pending_ == 0; });
        job_ = nullptr;
    }
private:
    void loop(int i) {
        int seen = 0;
        for (;;) {
            const std::function<void(int)>* job;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                job = job_;
            }
            (*job)(i);
            { std::lock_guard<std::mutex> l(m_); if (--pending_ == 0) done_.notify_one(); }
        }
    }
    int n_; std::vector<std::thread> th_; std::mutex m_; std::condition_variable cv_, done_;
    const std::function<void(int)>* job_ = nullptr; int pending_ = 0, gen_ = 0; bool stop_ = false;
};

// ------------------------------------------------------------------------------------------------
// When to narrow.  The host thread narrows chunk c+1 while the copy engine moves chunk c, so a chunk period is
// max(narrow time, int16 copy time) with narrowing and the float32 copy time without: narrowing pays exactly when
// the conversion rate (float32 bytes/s) beats the wire rate of the host->device copies.  Both are MEASURED by this
// decoder, under whatever the other ranks of the box are doing at the same time: conversion times per chunk on the
// host clock, copy times per chunk with CUDA events on the copy stream.  The verdict is taken once kNarrowVotes
// chunks of >= 1 M samples have been seen, from the MEDIANS of the last (at most 64) samples: on only if the conversion
// beats the copy by 25 %, and off again -- for good -- when a later call finds the win below 5 % (mode LOE_NARROW_AUTO);
// loe_decoder_set_narrow / LOE_B200_NARROW = on|off make it an explicit choice.
// ------------------------------------------------------------------------------------------------
constexpr int kNarrowVotes = 6;
constexpr double kNarrowOnMargin = 1.25;       // switch on (stay undecided -> on) only with a clear win ...
constexpr double kNarrowOffMargin = 1.05;      // ... and off again as soon as the win is gone: the verdict is re-examined on every
                                               // call from the medians of the last samples, so one lucky first call cannot latch it

// CPUs next to the GPU: /sys/bus/pci/devices/<bus id>/local_cpulist ("0-15,32-47"), intersected with the CPUs this
// process may use.  Empty when the file is unreadable (containers) -- the workers then float.
static std::vector<int> local_cpus(int device) {
    std::vector<int> out;
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) return out;
    for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE* f = fopen(path, "r");
    if (!f) return out;
    char line[4096] = "";
    if (!fgets(line, sizeof(line), f)) line[0] = 0;
    fclose(f);
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    const bool have_allowed = sched_getaffinity(0, sizeof(allowed), &allowed) == 0;
    for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k < 1) continue;
        if (k == 1) b = a;
        for (int c = a; c <= b && c < CPU_SETSIZE; ++c)
            if (!have_allowed || CPU_ISSET(c, &allowed)) out.push_back(c);
    }
    return out;
}

static int env_int(const char* name, int fallback) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : fallback;
}

static double median(std::vector<double> v) {
    if (v.empty()) return 0.0;
    std::sort(v.begin(), v.end());
    return v[v.size() / 2];
}

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return LOE_OK;
        if (p) { LOE_CUDA(cudaFree(p)); p = nullptr; cap = 0; }
        const size_t want = bytes + bytes / 8 + 256;
        LOE_CUDA(cudaMalloc(&p, want));
        cap = want;
        return LOE_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Decoder {
    int device = 0;
    cudaStream_t copy = nullptr, comp = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    cudaEvent_t ev_copy_begin[2] = {nullptr, nullptr};                          // timing pair of a chunk's PCM copy
    size_t copy_bytes[2] = {0, 0};
    // model
    int mel_na = 0, mel_nb = 0, n_states = 0, n_pos = 0, n_ends = 0;
    int32_t* d_mel_bin = nullptr; float* d_mel_w = nullptr; float* d_b = nullptr; float* d_cst = nullptr;
    void* d_b_h16 = nullptr;            // optional 3xFP16 image (loe_decoder_set_h16): then the FP16 tensor-core kernel scores
    int32_t* d_tr_off = nullptr; int32_t* d_col = nullptr; float* d_band = nullptr; uint8_t* d_flags = nullptr;
    int32_t* d_word = nullptr; int32_t* d_word_lo = nullptr;
    // workspace
    DevBuf pcm[2], off[2], mel, feat, scores, path, umax, words, count, best, best_score, bp, aimg, inv2;
    int64_t* h_off[2] = {nullptr, nullptr}; size_t h_off_cap[2] = {0, 0};      // pinned staging: [pcm_off | frm_off]
    bool used[2] = {false, false};
    char* h_out = nullptr; size_t h_out_cap = 0;                                // pinned staging of the results
    // host-side narrowing of float32 PCM (see above): worker pool, pinned int16 staging per buffer set, and the
    // running verdict: -1 = not measured yet, 0 = off (conversion slower than the float32 copy it saves), 1 = on
    Pool* pool = nullptr;
    int16_t* h_stage[2] = {nullptr, nullptr}; size_t h_stage_cap[2] = {0, 0};
    int narrow_mode = LOE_NARROW_AUTO;                                          // explicit choice, or auto
    int narrow = -1;                                                            // auto verdict: -1 undecided (narrowing while measuring)
    int narrow_threads = 0;                                                     // workers in use (0 until the pool exists)
    int pinned_cpus = 0;                                                        // CPUs the workers are bound to (0: floating)
    std::vector<double> narrow_samples, copy_samples;                           // GB/s per chunk: float32 bytes converted / wire bytes copied
    double narrow_gbps = 0.0, copy_gbps = 0.0;                                  // medians of the above
    // last call
    int64_t last_wire_bytes = 0, last_pcm_bytes = 0; int last_chunks = 0, last_chunks_narrowed = 0;
};

template <typename T>
static int upload(T** dst, const T* src, size_t n) {
    LOE_CUDA(cudaMalloc((void**)dst, sizeof(T) * std::max<size_t>(n, 1)));
    if (n) LOE_CUDA(cudaMemcpy(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice));
    return LOE_OK;
}

static void destroy(Decoder* d) {
    if (!d) return;
    cudaSetDevice(d->device);
    if (d->comp) cudaStreamSynchronize(d->comp);
    if (d->copy) cudaStreamSynchronize(d->copy);
    for (int i = 0; i < 2; ++i) {
        d->pcm[i].release(); d->off[i].release();
        if (d->h_off[i]) cudaFreeHost(d->h_off[i]);
        if (d->ev_copy[i]) cudaEventDestroy(d->ev_copy[i]);
        if (d->ev_done[i]) cudaEventDestroy(d->ev_done[i]);
        if (d->ev_copy_begin[i]) cudaEventDestroy(d->ev_copy_begin[i]);
    }
    if (d->h_out) cudaFreeHost(d->h_out);
    for (int i = 0; i < 2; ++i) if (d->h_stage[i]) cudaFreeHost(d->h_stage[i]);
    delete d->pool;
    DevBuf* bufs[] = {&d->mel, &d->feat, &d->scores, &d->path, &d->umax, &d->words, &d->count, &d->best, &d->best_score, &d->bp, &d->aimg, &d->inv2};
    for (DevBuf* b : bufs) b->release();
    void* tabs[] = {d->d_b_h16, d->d_mel_bin, d->d_mel_w, d->d_b, d->d_cst, d->d_tr_off, d->d_col, d->d_band, d->d_flags, d->d_word, d->d_word_lo};
    for (void* t : tabs) if (t) cudaFree(t);
    if (d->copy) cudaStreamDestroy(d->copy);
    if (d->comp) cudaStreamDestroy(d->comp);
    delete d;
}

}  // namespace loe

extern "C" int loe_decoder_create(int device, const int32_t* mel_bin_host, const float* mel_w_host, int mel_na, int mel_nb,
                                  const float* b_packed_host, const float* cst_pad_host, int n_states,
                                  int n_pos, const int32_t* col_host, const float* band_host, const uint8_t* flags_host,
                                  const int32_t* word_host, const int32_t* word_lo_host, void** out) {
    using namespace loe;
    if (!out) { set_error("out is NULL"); return LOE_ERR_VALUE; }
    *out = nullptr;
    if (n_pos <= 0 || n_pos > LOE_MAX_POS) { set_error("%d trellis positions (1..%d supported)", n_pos, LOE_MAX_POS); return n_pos > LOE_MAX_POS ? LOE_ERR_OVERFLOW : LOE_ERR_VALUE; }
    if (n_states <= 0 || mel_na < 0 || mel_nb < 0 || mel_na > LOE_MEL_NA_MAX || mel_nb > LOE_MEL_NB_MAX) { set_error("bad model sizes"); return LOE_ERR_VALUE; }
    LOE_CUDA(cudaSetDevice(device));
    Decoder* d = new (std::nothrow) Decoder();
    if (!d) { set_error("out of host memory"); return LOE_ERR_CUDA; }
    d->device = device; d->mel_na = mel_na; d->mel_nb = mel_nb; d->n_states = n_states; d->n_pos = n_pos;
    if (const char* e = getenv("LOE_B200_NARROW")) {
        if (!strcmp(e, "on") || !strcmp(e, "1")) d->narrow_mode = LOE_NARROW_ON;
        else if (!strcmp(e, "off") || !strcmp(e, "0")) d->narrow_mode = LOE_NARROW_OFF;
    }
    for (int p = 0; p < n_pos; ++p) d->n_ends += (flags_host[p] & LOE_POS_END) ? 1 : 0;
    int st = LOE_OK;
    const int n_tiles = loe_emission_tc_tiles(n_states);
    const int32_t tr_off[2] = {0, n_pos};
    auto fail = [&](int code) { destroy(d); return code; };
#define LOE_TRY(x) if ((st = (x)) != LOE_OK) return fail(st)
    LOE_TRY(check_cuda(cudaStreamCreateWithFlags(&d->copy, cudaStreamNonBlocking), "stream"));
    LOE_TRY(check_cuda(cudaStreamCreateWithFlags(&d->comp, cudaStreamNonBlocking), "stream"));
    for (int i = 0; i < 2; ++i) {
        LOE_TRY(check_cuda(cudaEventCreate(&d->ev_copy[i]), "event"));
        LOE_TRY(check_cuda(cudaEventCreate(&d->ev_copy_begin[i]), "event"));
        LOE_TRY(check_cuda(cudaEventCreateWithFlags(&d->ev_done[i], cudaEventDisableTiming), "event"));
    }
    LOE_TRY(upload(&d->d_mel_bin, mel_bin_host, (size_t)(mel_na + mel_nb) * 32));
    LOE_TRY(upload(&d->d_mel_w, mel_w_host, (size_t)(mel_na + mel_nb) * 32));
    LOE_TRY(upload(&d->d_b, b_packed_host, (size_t)n_tiles * 19200));
    LOE_TRY(upload(&d->d_cst, cst_pad_host, (size_t)n_tiles * 6));
    LOE_TRY(upload(&d->d_tr_off, tr_off, 2));
    LOE_TRY(upload(&d->d_col, col_host, (size_t)n_pos));
    LOE_TRY(upload(&d->d_band, band_host, (size_t)n_pos * 3));
    LOE_TRY(upload(&d->d_flags, flags_host, (size_t)n_pos));
    LOE_TRY(upload(&d->d_word, word_host, (size_t)n_pos));
    LOE_TRY(upload(&d->d_word_lo, word_lo_host, (size_t)n_pos));
#undef LOE_TRY
    *out = d;
    return LOE_OK;
}

extern "C" int loe_decoder_set_h16(void* dec, const void* b_h16_host) {
    using namespace loe;
    Decoder* d = reinterpret_cast<Decoder*>(dec);
    if (!d) { set_error("decoder is NULL"); return LOE_ERR_VALUE; }
    LOE_CUDA(cudaSetDevice(d->device));
    if (d->d_b_h16) { cudaFree(d->d_b_h16); d->d_b_h16 = nullptr; }
    if (!b_h16_host) return LOE_OK;
    const size_t bytes = (size_t)loe_emission_tc_tiles(d->n_states) * (size_t)loe_emission_h16_tile_bytes();
    LOE_CUDA(cudaMalloc(&d->d_b_h16, bytes));
    LOE_CUDA(cudaMemcpy(d->d_b_h16, b_h16_host, bytes, cudaMemcpyHostToDevice));
    return LOE_OK;
}

extern "C" void loe_decoder_destroy(void* dec) { loe::destroy(reinterpret_cast<loe::Decoder*>(dec)); }

extern "C" int loe_decoder_decode_host(void* dec, const void* pcm_host, int pcm_format, const int64_t* sample_off_host, int n_utt,
                                       double penalty, int penalty_f64, int skip_label, int max_words, int n_chunks,
                                       int8_t* words_host, int32_t* count_host, float* best_score_host, int8_t* path_host) {
    using namespace loe;
    Decoder* d = reinterpret_cast<Decoder*>(dec);
    if (!d) { set_error("decoder is NULL"); return LOE_ERR_VALUE; }
    if (n_utt <= 0) return LOE_OK;
    if (pcm_format != LOE_PCM_F32 && pcm_format != LOE_PCM_S16) { set_error("unknown pcm_format %d", pcm_format); return LOE_ERR_VALUE; }
    if (max_words <= 0 || !words_host || !count_host) { set_error("words / count buffers required"); return LOE_ERR_VALUE; }
    LOE_CUDA(cudaSetDevice(d->device));
    const size_t bps = pcm_format == LOE_PCM_F32 ? 4 : 2;
    const int64_t total_samples = sample_off_host[n_utt] - sample_off_host[0];
    // default: chunks of about 64 MB (the first chunk's copy and the last chunk's kernels are the part of the call
    // that nothing overlaps), at most 64
    if (n_chunks <= 0) n_chunks = (int)std::min<int64_t>(64, std::max<int64_t>(1, (int64_t)(total_samples * bps) / (64ll << 20)));
    // chunk boundaries: whole utterances, about equal numbers of samples
    std::vector<int> bounds{0};
    for (int c = 1; c < n_chunks; ++c) {
        const int64_t target = sample_off_host[0] + total_samples * c / n_chunks;
        int u = (int)(std::lower_bound(sample_off_host, sample_off_host + n_utt + 1, target) - sample_off_host);
        u = std::min(std::max(u, bounds.back()), n_utt);
        if (u > bounds.back()) bounds.push_back(u);
    }
    if (bounds.back() != n_utt) bounds.push_back(n_utt);
    // results land in pinned staging (a device->host copy into pageable memory would block the host
    // until the chunk's kernels finish and with it the next chunk's upload), then move to the caller's arrays
    int64_t total_frames = 0;
    for (int i = 0; i < n_utt; ++i) total_frames += 1 + (sample_off_host[i + 1] - sample_off_host[i]) / LOE_HOP;
    const size_t o_words = 0, o_count = o_words + (((size_t)n_utt * max_words + 15) & ~(size_t)15),
                 o_score = o_count + (size_t)n_utt * 4, o_path = o_score + (size_t)n_utt * 4,
                 out_bytes = o_path + (path_host ? (size_t)total_frames : 0);
    if (d->h_out_cap < out_bytes) {
        if (d->h_out) LOE_CUDA(cudaFreeHost(d->h_out));
        d->h_out = nullptr; d->h_out_cap = 0;
        LOE_CUDA(cudaHostAlloc((void**)&d->h_out, out_bytes + out_bytes / 4, cudaHostAllocDefault));
        d->h_out_cap = out_bytes + out_bytes / 4;
    }
    // worker threads of the float32 -> int16 narrowing: this rank's share of the machine -- hardware threads divided
    // by the ranks on the box (LOCAL_WORLD_SIZE, as torchrun exports it), at most 32 -- bound to the CPUs next to the
    // GPU when the kernel tells which those are.  LOE_B200_NARROW_THREADS overrides the count (0 = never narrow),
    // LOE_B200_NARROW_PIN=0 leaves the workers floating.
    const bool narrowing_possible = pcm_format == LOE_PCM_F32 && d->narrow_mode != LOE_NARROW_OFF &&
                                    !(d->narrow_mode == LOE_NARROW_AUTO && d->narrow == 0);
    int narrow_threads = 0;
    if (narrowing_possible) {
        const int hw = (int)std::max<unsigned>(1u, std::thread::hardware_concurrency());
        const int ranks = std::max(1, env_int("LOCAL_WORLD_SIZE", 1));
        narrow_threads = std::max(1, std::min(32, hw / ranks));
        if (const char* e = getenv("LOE_B200_NARROW_THREADS")) narrow_threads = std::max(0, std::min(256, atoi(e)));
        if (d->pool && d->pool->size() != narrow_threads) { delete d->pool; d->pool = nullptr; }
        if (!d->pool && narrow_threads > 0) {
            std::vector<int> cpus;
            if (env_int("LOE_B200_NARROW_PIN", 1) != 0) cpus = local_cpus(d->device);
            if ((int)cpus.size() < narrow_threads) cpus.clear();        // fewer local CPUs than workers: let them float
            d->pool = new (std::nothrow) Pool(narrow_threads, cpus);
            d->pinned_cpus = d->pool ? (int)cpus.size() : 0;
        }
        d->narrow_threads = d->pool ? narrow_threads : 0;
    }
    d->last_wire_bytes = 0; d->last_pcm_bytes = (int64_t)total_samples * (int64_t)bps; d->last_chunks = 0; d->last_chunks_narrowed = 0;
    // conversion and copy rates are sampled on every chunk of at least min_samples samples (reported by
    // loe_decoder_stats); the automatic verdict is taken from them while it is still open
    const int64_t min_samples = env_int("LOE_B200_NARROW_MIN_SAMPLES", 1 << 20);
    const bool deciding = pcm_format == LOE_PCM_F32 && d->narrow_mode == LOE_NARROW_AUTO && d->narrow < 0;
    auto keep = [](std::vector<double>& v, double x) { if (v.size() >= 64) v.erase(v.begin(), v.begin() + 32); v.push_back(x); };
    int64_t frames_done = 0;
    for (size_t c = 0; c + 1 < bounds.size(); ++c) {
        const int a = bounds[c], b = bounds[c + 1], n = b - a, set = (int)(c & 1);
        const int64_t s0 = sample_off_host[a], ns = sample_off_host[b] - s0;
        // staging of this set is free once its previous copy has completed
        // (host) and its device buffers once the kernels of the chunk that used them are done (copy stream waits)
        if (d->used[set]) { LOE_CUDA(cudaEventSynchronize(d->ev_copy[set])); LOE_CUDA(cudaStreamWaitEvent(d->copy, d->ev_done[set], 0)); }
        const size_t off_elems = 2 * (size_t)(n + 1);
        if (d->h_off_cap[set] < off_elems) {
            if (d->h_off[set]) LOE_CUDA(cudaFreeHost(d->h_off[set]));
            d->h_off[set] = nullptr; d->h_off_cap[set] = 0;
            LOE_CUDA(cudaHostAlloc((void**)&d->h_off[set], sizeof(int64_t) * (off_elems + off_elems / 4 + 16), cudaHostAllocDefault));
            d->h_off_cap[set] = off_elems + off_elems / 4 + 16;
        }
        int64_t* pcm_off = d->h_off[set];
        int64_t* frm_off = pcm_off + (n + 1);
        int max_frames = 0, min_frames = 1 << 30;
        frm_off[0] = 0;
        for (int i = 0; i <= n; ++i) pcm_off[i] = sample_off_host[a + i] - s0;
        for (int i = 0; i < n; ++i) {
            const int fr = (int)(1 + (pcm_off[i + 1] - pcm_off[i]) / LOE_HOP);
            frm_off[i + 1] = frm_off[i] + fr;
            max_frames = std::max(max_frames, fr); min_frames = std::min(min_frames, fr);
        }
        const int64_t F = frm_off[n];
        int st;
        // float32 PCM: try to send the chunk as int16 (exact or not at all); the conversion of this chunk overlaps the
        // copy and the kernels of the previous one
        const char* src = (const char*)pcm_host + (size_t)(s0 - sample_off_host[0]) * bps;
        int chunk_format = pcm_format;
        size_t chunk_bps = bps;
        if (narrowing_possible && d->pool && ns > 0) {
            if (d->h_stage_cap[set] < (size_t)ns) {
                if (d->h_stage[set]) LOE_CUDA(cudaFreeHost(d->h_stage[set]));
                d->h_stage[set] = nullptr; d->h_stage_cap[set] = 0;
                const size_t want = (size_t)ns + (size_t)ns / 8 + 64;
                LOE_CUDA(cudaHostAlloc((void**)&d->h_stage[set], want * sizeof(int16_t), cudaHostAllocDefault));
                d->h_stage_cap[set] = want;
            }
            const float* fsrc = reinterpret_cast<const float*>(src);
            int16_t* fdst = d->h_stage[set];
            std::atomic<int> exact(1);
            const int W = d->pool->size();
            const auto t0 = std::chrono::steady_clock::now();
            d->pool->run([&](int w) {
                const int64_t i0 = (ns * w / W) & ~(int64_t)15, i1 = (w == W - 1) ? ns : ((ns * (w + 1) / W) & ~(int64_t)15);
                if (i1 > i0 && !narrow_range(fsrc + i0, fdst + i0, i1 - i0)) exact.store(0);
            });
            const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (ns >= min_samples && sec > 0) keep(d->narrow_samples, (double)ns * 4.0 / sec * 1e-9);
            if (exact.load()) { src = reinterpret_cast<const char*>(fdst); chunk_format = LOE_PCM_S16; chunk_bps = 2; d->last_chunks_narrowed++; }
        }
        if ((st = d->pcm[set].ensure((size_t)ns * chunk_bps)) != LOE_OK) return st;
        if ((st = d->off[set].ensure(sizeof(int64_t) * off_elems)) != LOE_OK) return st;
        LOE_CUDA(cudaMemcpyAsync(d->off[set].p, pcm_off, sizeof(int64_t) * off_elems, cudaMemcpyHostToDevice, d->copy));
        if (d->copy_bytes[set]) {                  // the previous copy of this set has completed (synchronised above)
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, d->ev_copy_begin[set], d->ev_copy[set]) == cudaSuccess && ms > 0.f)
                keep(d->copy_samples, (double)d->copy_bytes[set] / (ms * 1e-3) * 1e-9);
            d->copy_bytes[set] = 0;
        }
        LOE_CUDA(cudaEventRecord(d->ev_copy_begin[set], d->copy));
        LOE_CUDA(cudaMemcpyAsync(d->pcm[set].p, src, (size_t)ns * chunk_bps, cudaMemcpyHostToDevice, d->copy));
        LOE_CUDA(cudaEventRecord(d->ev_copy[set], d->copy));
        if (ns >= min_samples) d->copy_bytes[set] = (size_t)ns * chunk_bps;
        d->last_wire_bytes += (int64_t)((size_t)ns * chunk_bps + sizeof(int64_t) * off_elems);
        d->last_chunks++;
        d->used[set] = true;
        // compute buffers are shared by all chunks: growing them must wait for the chunks in flight
        const bool bp_needed = !loe_viterbi_bp_fits(max_frames, d->n_pos);
        // with the 3xFP16 image the cepstrum kernel writes the emission kernel's A operand directly (no float32 feature
        // matrix at all): see loe_mfcc_img_dev / loe_emission_h16_img_dev
        const bool img = d->d_b_h16 != nullptr;
        const size_t need[] = {(size_t)F * 40 * 4, img ? 0 : (size_t)F * 39 * 4, (size_t)F * d->n_states * 4, (size_t)F, (size_t)n * 4,
                               (size_t)n * max_words, (size_t)n * 4, (size_t)n * 4, (size_t)n * 4, bp_needed ? (size_t)F * LOE_MAX_POS : 0,
                               img ? (size_t)loe_emission_h16_img_bytes(F) : 0, img ? (size_t)((F + 127) / 128) * 128 * 4 : 0};
        DevBuf* bufs[] = {&d->mel, &d->feat, &d->scores, &d->path, &d->umax, &d->words, &d->count, &d->best, &d->best_score, &d->bp,
                          &d->aimg, &d->inv2};
        bool grow = false;
        for (int i = 0; i < 12; ++i) grow |= need[i] > bufs[i]->cap;
        if (grow) {
            LOE_CUDA(cudaStreamSynchronize(d->comp));
            for (int i = 0; i < 12; ++i) if ((st = bufs[i]->ensure(need[i])) != LOE_OK) return st;
        }
        LOE_CUDA(cudaStreamWaitEvent(d->comp, d->ev_copy[set], 0));
        const int64_t* d_pcm_off = (const int64_t*)d->off[set].p;
        const int64_t* d_frm_off = d_pcm_off + (n + 1);
        if (img) {
            if ((st = loe_mfcc_img_dev(d->pcm[set].p, chunk_format, d_pcm_off, d_frm_off, n, F, max_frames, min_frames, d->d_mel_bin, d->d_mel_w,
                                       d->mel_na, d->mel_nb, (float*)d->mel.p, (float*)d->umax.p, nullptr, d->aimg.p, (float*)d->inv2.p,
                                       d->comp, 3)) != LOE_OK) return st;
            st = loe_emission_h16_img_dev(d->aimg.p, (const float*)d->inv2.p, F, d->d_b_h16, d->d_cst, d->n_states, (float*)d->scores.p,
                                          d->n_states, d->comp);
        } else {
            if ((st = loe_mfcc_dev(d->pcm[set].p, chunk_format, d_pcm_off, d_frm_off, n, F, max_frames, min_frames, d->d_mel_bin, d->d_mel_w,
                                   d->mel_na, d->mel_nb, (float*)d->mel.p, (float*)d->umax.p, (float*)d->feat.p, d->comp)) != LOE_OK) return st;
            st = loe_emission_tc_dev((const float*)d->feat.p, F, 39, d->d_b, d->d_cst, d->n_states, (float*)d->scores.p, d->n_states, d->comp);
        }
        if (st != LOE_OK) return st;
        if ((st = loe_viterbi_dev((const float*)d->scores.p, d->n_states, d_frm_off, n, max_frames, d->d_tr_off, d->d_col, d->d_band,
                                  d->d_flags, d->n_pos, nullptr, 1, penalty, penalty_f64, (int8_t*)d->path.p, nullptr, d->n_ends,
                                  (int32_t*)d->best.p, (float*)d->best_score.p, bp_needed ? (uint8_t*)d->bp.p : nullptr,
                                  d->d_word, d->d_word_lo, skip_label, (int8_t*)d->words.p, max_words, (int32_t*)d->count.p, d->comp)) != LOE_OK)
            return st;
        LOE_CUDA(cudaMemcpyAsync(d->h_out + o_words + (size_t)a * max_words, d->words.p, (size_t)n * max_words, cudaMemcpyDeviceToHost, d->comp));
        LOE_CUDA(cudaMemcpyAsync(d->h_out + o_count + (size_t)a * 4, d->count.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, d->comp));
        if (best_score_host) LOE_CUDA(cudaMemcpyAsync(d->h_out + o_score + (size_t)a * 4, d->best_score.p, sizeof(float) * n, cudaMemcpyDeviceToHost, d->comp));
        if (path_host) LOE_CUDA(cudaMemcpyAsync(d->h_out + o_path + frames_done, d->path.p, (size_t)F, cudaMemcpyDeviceToHost, d->comp));
        LOE_CUDA(cudaEventRecord(d->ev_done[set], d->comp));
        frames_done += F;
    }
    LOE_CUDA(cudaStreamSynchronize(d->comp));
    for (int set = 0; set < 2; ++set) {
        if (!d->copy_bytes[set]) continue;
        float ms = 0.f;
        if (cudaEventSynchronize(d->ev_copy[set]) == cudaSuccess &&
            cudaEventElapsedTime(&ms, d->ev_copy_begin[set], d->ev_copy[set]) == cudaSuccess && ms > 0.f)
            keep(d->copy_samples, (double)d->copy_bytes[set] / (ms * 1e-3) * 1e-9);
        d->copy_bytes[set] = 0;
    }
    d->narrow_gbps = median(d->narrow_samples);
    d->copy_gbps = median(d->copy_samples);
    if (pcm_format == LOE_PCM_F32 && d->narrow_mode == LOE_NARROW_AUTO && d->narrow != 0 &&
        (int)d->narrow_samples.size() >= kNarrowVotes && (int)d->copy_samples.size() >= kNarrowVotes) {
        if (d->narrow < 0) d->narrow = d->narrow_gbps > kNarrowOnMargin * d->copy_gbps ? 1 : 0;
        else if (d->narrow_gbps < kNarrowOffMargin * d->copy_gbps) d->narrow = 0;        // once off it stays off (no more samples)
    }
    (void)deciding;
    memcpy(words_host, d->h_out + o_words, (size_t)n_utt * max_words);
    memcpy(count_host, d->h_out + o_count, (size_t)n_utt * 4);
    if (best_score_host) memcpy(best_score_host, d->h_out + o_score, (size_t)n_utt * 4);
    if (path_host) memcpy(path_host, d->h_out + o_path, (size_t)total_frames);
    return LOE_OK;
}

extern "C" int loe_pcm_narrow_host(const float* src_host, int16_t* dst_host, int64_t n_samples) {
    if (n_samples <= 0) return 1;
    if (!src_host || !dst_host) return 0;
    return loe::narrow_range(src_host, dst_host, n_samples) ? 1 : 0;
}

extern "C" double loe_decoder_narrow_rate(void* dec) {
    loe::Decoder* d = reinterpret_cast<loe::Decoder*>(dec);
    if (!d) return 0.0;
    const bool off = d->narrow_mode == LOE_NARROW_OFF || (d->narrow_mode == LOE_NARROW_AUTO && d->narrow == 0);
    return off ? -d->narrow_gbps : d->narrow_gbps;
}

extern "C" int loe_decoder_set_narrow(void* dec, int mode) {
    using namespace loe;
    Decoder* d = reinterpret_cast<Decoder*>(dec);
    if (!d) { set_error("decoder is NULL"); return LOE_ERR_VALUE; }
    if (mode != LOE_NARROW_AUTO && mode != LOE_NARROW_OFF && mode != LOE_NARROW_ON) { set_error("unknown narrow mode %d", mode); return LOE_ERR_VALUE; }
    d->narrow_mode = mode;
    if (mode == LOE_NARROW_AUTO) { d->narrow = -1; d->narrow_samples.clear(); d->copy_samples.clear(); d->narrow_gbps = d->copy_gbps = 0.0; }
    return LOE_OK;
}

extern "C" int loe_decoder_stats(void* dec, double* out, int n) {
    using namespace loe;
    Decoder* d = reinterpret_cast<Decoder*>(dec);
    if (!d || !out) { set_error("decoder / out is NULL"); return LOE_ERR_VALUE; }
    const int on = d->narrow_mode == LOE_NARROW_ON ? 1 : d->narrow_mode == LOE_NARROW_OFF ? 0 : d->narrow;   // -1: still measuring
    const double v[LOE_DECODER_STATS] = {(double)d->narrow_mode, (double)on, d->narrow_gbps, d->copy_gbps, (double)d->narrow_threads,
                                         (double)d->pinned_cpus, (double)d->last_pcm_bytes, (double)d->last_wire_bytes,
                                         (double)d->last_chunks, (double)d->last_chunks_narrowed};
    for (int i = 0; i < n && i < LOE_DECODER_STATS; ++i) out[i] = v[i];
    return LOE_OK;
}

// Content fingerprint of a model's host arrays (the pack cache of the Python layer asks "were these edited in place?"
// before every single-utterance call).  Every 8-byte word goes through h = rotl(h ^ w, 23) + k in one of 16 lanes (AVX2:
// four 256-bit accumulators) -- a bijection of the lane for a given word and of the word for a given lane, so a change of
// any single word always changes the result; tail bytes and block lengths are folded in with multiply-xor steps.  Host
// only, no CUDA call.
namespace loe {
static inline uint64_t fp_mix(uint64_t h, uint64_t w) {
    h ^= w;
    return ((h << 23) | (h >> 41)) + 0x9E3779B97F4A7C15ull;
}
static void fp_words_scalar(uint64_t* h, const unsigned char* p, int64_t n_chunks) {       // chunks of 128 bytes
    for (int64_t c = 0; c < n_chunks; ++c, p += 128) {
        uint64_t w[16];
        memcpy(w, p, 128);
        for (int l = 0; l < 16; ++l) h[l] = fp_mix(h[l], w[l]);
    }
}
__attribute__((target("avx2"))) static void fp_words_avx2(uint64_t* h, const unsigned char* p, int64_t n_chunks) {
    __m256i a[4];
    for (int v = 0; v < 4; ++v) a[v] = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(h + 4 * v));
    const __m256i k = _mm256_set1_epi64x((long long)0x9E3779B97F4A7C15ull);
    for (int64_t c = 0; c < n_chunks; ++c, p += 128)
        for (int v = 0; v < 4; ++v) {
            const __m256i x = _mm256_xor_si256(a[v], _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + 32 * v)));
            a[v] = _mm256_add_epi64(_mm256_or_si256(_mm256_slli_epi64(x, 23), _mm256_srli_epi64(x, 41)), k);
        }
    for (int v = 0; v < 4; ++v) _mm256_storeu_si256(reinterpret_cast<__m256i*>(h + 4 * v), a[v]);
}
}  // namespace loe

extern "C" uint64_t loe_host_fingerprint(const void* const* blocks, const int64_t* n_bytes, int n_blocks) {
    using namespace loe;
    static const bool avx2 = __builtin_cpu_supports("avx2");
    const uint64_t k = 0x9E3779B97F4A7C15ull;
    uint64_t h[16];
    for (int l = 0; l < 16; ++l) h[l] = 0x243F6A8885A308D3ull * (uint64_t)(2 * l + 1);
    for (int b = 0; b < n_blocks; ++b) {
        const unsigned char* p = static_cast<const unsigned char*>(blocks[b]);
        const int64_t n = n_bytes[b];
        if (!p || n <= 0) { h[0] = (h[0] ^ (uint64_t)(n + 1)) * k; continue; }
        const int64_t chunks = n / 128;
        if (avx2) fp_words_avx2(h, p, chunks); else fp_words_scalar(h, p, chunks);
        uint64_t tail[16] = {0};
        memcpy(tail, p + chunks * 128, (size_t)(n - chunks * 128));
        for (int l = 0; l < 16; ++l) h[l] = fp_mix(h[l], tail[l]);
        h[b & 15] = (h[b & 15] ^ (uint64_t)n) * k;
    }
    uint64_t r = 0;
    for (int l = 0; l < 16; ++l) { r = (r ^ h[l]) * k; r ^= r >> 32; }
    return r;
}

// Word-id tables -> text (host code): utterance i contributes the single-character labels of its first count[i] word ids
// followed by ``sep``; an utterance whose count does not fit the table (count < 0: T == 1, or count > max_words)
// contributes ``sep`` alone -- the caller decodes those from the state path.  Returns the number of bytes written
// (at most n_utt * (max_words + 1)).
extern "C" int64_t loe_labels_text_host(const int8_t* words_host, const int32_t* count_host, int n_utt, int max_words,
                                        const char* label_chars, int n_labels, char sep, char* out_host) {
    if (!words_host || !count_host || !label_chars || !out_host || n_utt <= 0 || max_words <= 0 || n_labels <= 0) return 0;
    char* o = out_host;
    for (int i = 0; i < n_utt; ++i) {
        const int c = count_host[i];
        if (c >= 0 && c <= max_words) {
            const int8_t* w = words_host + (size_t)i * max_words;
            for (int k = 0; k < c; ++k) {
                int id = w[k];
                id = id < 0 ? 0 : id >= n_labels ? n_labels - 1 : id;
                *o++ = label_chars[id];
            }
        }
        *o++ = sep;
    }
    return (int64_t)(o - out_host);
}

extern "C" int loe_host_alloc(void** ptr_out, size_t bytes) {
    using namespace loe;
    if (!ptr_out) { set_error("ptr_out is NULL"); return LOE_ERR_VALUE; }
    LOE_CUDA(cudaHostAlloc(ptr_out, bytes ? bytes : 1, cudaHostAllocDefault));
    return LOE_OK;
}

extern "C" int loe_host_free(void* ptr) {
    using namespace loe;
    if (ptr) LOE_CUDA(cudaFreeHost(ptr));
    return LOE_OK;
}
