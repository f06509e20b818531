// Parameterised MFCC front end for sm_100a (BASELINE.json configs[3], "spec" parameter set: 25 ms / 10 ms frames,
// 512-point FFT, Hamming window, pre-emphasis 0.97, 40 mel bands, natural log, 13 cepstra + delta + delta-delta, CMN).
//
// The reference has ONE parameter set (mfcc.py:31-34: n_fft 320, hop 160, Hann, dB, per-frame normalisation), served by
// the specialised kernels of mfcc.cu.  This file is the same librosa-shaped pipeline (mfcc.py:24-44) with the stages
// the north star names made parameters: any power-of-two FFT size up to 1024, any window (host table, so Hamming /
// Hann / a short window centred in a longer FFT), pre-emphasis, dB or natural log, and the normalisation of the static
// block (none / per frame as the reference does / cepstral mean / mean and variance over the utterance).
//
//   mel_ex_kernel    PCM -> pre-emphasis -> window -> N-point real FFT (N/2-point complex Stockham radix-4 FFT of one
//                    frame per warp in shared memory, padded against bank conflicts, + real-input post-pass) -> |.|^2
//                    -> triangular filterbank as a lane-balanced table -> mel energies [frames, n_mels]
//                    (+ per-utterance maximum for the dB mode)
//   ceps_ex_kernel   mel -> log -> DCT-II -> cepstra [frames, n_ceps]            (thread per frame)
//   cmn_ex_kernel    per-utterance mean (and 1/std) of every cepstral coefficient, fixed summation order (CTA per utterance)
//   feat_ex_kernel   normalised static block + Savitzky-Golay delta / delta-delta (width 9) -> features [frames, 3 n_ceps]
//
// Algorithmic HBM bytes per frame: 4 * hop (PCM) + 12 * n_ceps (features).
#include "common.cuh"
#include "fft_regs.cuh"
#include <math.h>
#include <atomic>
#include <mutex>

namespace loe {

constexpr int kWarpsE = 4;
constexpr int kMaxMelsE = 64;
constexpr int kMaxCepsE = 16;

__device__ __forceinline__ float2 cmulE(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float sample_f32(const float* x, int64_t i) { return __ldg(x + i); }
__device__ __forceinline__ float sample_f32(const short* x, int64_t i) { return (float)__ldg(x + i); }

// Shared memory of mel_ex_kernel<LOG2N>: per warp two ping-pong buffers of M = N/2 complex points (the second one
// also takes the power spectrum), CTA-wide the twiddles W_M^j (j < M), W_N^k (k <= M), the window, and the filterbank
// as a lane-balanced table (built by the CTA from the per-filter rows the caller passes).
// Buffer index i lives at pad(i) = i + (i >> 2): the strided stores of the first two Stockham passes (stride 4 and
// blocks of 4 at stride 16) then hit 32 different banks per half-warp instead of 8.
__device__ __forceinline__ int padE(int i) { return i + (i >> 2); }

template <int LOG2N>
struct SmemE {
    static constexpr int N = 1 << LOG2N, M = N / 2, MP = M + M / 4 + 4;
    float2 buf[kWarpsE][2][MP];
    float2 wm[M];
    float2 wn[M + 1];
    float win[N];
    int binA[32], binB[32];
    int na, nb, lb;
};

template <typename SampleT> struct PairE;
template <> struct PairE<float> { using type = float2; };
template <> struct PairE<short> { using type = short2; };

template <typename SampleT, int LOG2N>
__global__ void __launch_bounds__(kWarpsE * 32)
mel_ex_kernel(const SampleT* __restrict__ pcm, const int64_t* __restrict__ pcm_off, const int64_t* __restrict__ frm_off,
              const float* __restrict__ window, int hop, float preemph,
              const int32_t* __restrict__ mel_start, const int32_t* __restrict__ mel_len, const float* __restrict__ mel_w,
              int mel_pitch, int n_mels, int chunk, float* __restrict__ mel_out, float* __restrict__ utt_max) {
    using S = SmemE<LOG2N>;
    constexpr int N = S::N, M = S::M;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw_e[];
    S& sm = *reinterpret_cast<S*>(smem_raw_e);
    float* s_w = reinterpret_cast<float*>(smem_raw_e + sizeof(S));        // [(na + nb) * 32] lane table of filter weights
    const int u = blockIdx.x;
    const int64_t f0 = frm_off[u];
    const int T = (int)(frm_off[u + 1] - f0);
    const int t_begin = blockIdx.y * chunk;
    if (t_begin >= T) return;
    const int t_end = min(T, t_begin + chunk);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = tid; j < M; j += kWarpsE * 32) {
        float s, c;
        sincospif(-2.0f * (float)j / (float)M, &s, &c);
        sm.wm[j] = make_float2(c, s);
    }
    for (int k = tid; k <= M; k += kWarpsE * 32) {
        float s, c;
        sincospif(-2.0f * (float)k / (float)N, &s, &c);
        sm.wn[k] = make_float2(c, s);
    }
    for (int i = tid; i < N; i += kWarpsE * 32) sm.win[i] = window[i];
    // Filterbank as a lane table.  Round A: lane l owns filter l (l < 32), one bin per iteration, na = widest of them.
    // Round B: the filters from 32 on share the warp, lb = 32 / (their count rounded up to a power of two) lanes each: lane
    // q * lb + j walks the bins start + j, start + j + lb, ...; nb iterations; partial sums meet in xor shuffles.
    const int nA = min(n_mels, 32), nB = n_mels - nA;
    if (warp == 0) {
        int la = lane < nA ? mel_len[lane] : 0;
        int lbn = lane < nB ? mel_len[32 + lane] : 0;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) { la = max(la, __shfl_xor_sync(FULL, la, o)); lbn = max(lbn, __shfl_xor_sync(FULL, lbn, o)); }
        int lb = 32;
        while (lb > 1 && lb * nB > 32) lb >>= 1;
        if (lane == 0) { sm.na = la; sm.lb = lb; sm.nb = nB ? (lbn + lb - 1) / lb : 0; }
    }
    __syncthreads();
    const int na = sm.na, nb = sm.nb, lb = sm.lb;
    for (int i = tid; i < na * 32; i += kWarpsE * 32) {
        const int it = i >> 5, l = i & 31;
        s_w[i] = (l < nA && it < mel_len[l]) ? mel_w[(size_t)l * mel_pitch + it] : 0.f;
    }
    for (int i = tid; i < nb * 32; i += kWarpsE * 32) {
        const int it = i >> 5, l = i & 31, q = l / lb, j = l % lb, m = 32 + q, e = j + lb * it;
        s_w[na * 32 + i] = (q < nB && e < mel_len[m]) ? mel_w[(size_t)m * mel_pitch + e] : 0.f;
    }
    if (tid < 32) {
        sm.binA[tid] = tid < nA ? mel_start[tid] : 0;
        const int q = tid / lb;
        sm.binB[tid] = q < nB ? mel_start[32 + q] + tid % lb : 0;
    }
    __syncthreads();
    const int binA = sm.binA[lane], binB = sm.binB[lane];

    const int64_t s0 = pcm_off[u];
    const int64_t n_samples = pcm_off[u + 1] - s0;
    const SampleT* __restrict__ x = pcm + s0;
    using PairT = typename PairE<SampleT>::type;
    // two samples per load when every frame of this utterance starts on an even sample (base = hop t - N/2 + s0)
    const bool pair_ok = ((s0 & 1) == 0) && ((hop & 1) == 0) && ((reinterpret_cast<uintptr_t>(pcm) & (2 * sizeof(SampleT) - 1)) == 0);
    float2* a = sm.buf[warp][0];
    float2* b = sm.buf[warp][1];
    float vmax = 0.f;
    // emphasised sample i of the utterance (zero outside: centre padding); y[0] keeps (1 - preemph) x[0]
    auto emph = [&](int64_t i) -> float {
        if (i < 0 || i >= n_samples) return 0.f;
        const float cur = sample_f32(x, i);
        if (preemph == 0.f) return cur;
        const float prev = sample_f32(x, i > 0 ? i - 1 : 0);
        return __fsub_rn(cur, __fmul_rn(preemph, prev));         // never contracted: float32 and int16 PCM agree bit for bit
    };
    for (int t = t_begin + warp; t < t_end; t += kWarpsE) {
        const int64_t base = (int64_t)hop * t - N / 2;
        // z[n] = (w[2n] y[2n], w[2n+1] y[2n+1])
        if (pair_ok && base >= 1 && base + N <= n_samples) {
            // interior frame: one 2-sample load per point; the sample before a lane's pair is the previous lane's second
            // sample (lane 0: the last lane's of the round before, or one extra load)
            float carry = sample_f32(x, base - 1);
#pragma unroll 4
            for (int n = lane; n < M; n += 32) {
                const PairT p = __ldg(reinterpret_cast<const PairT*>(x + base + 2 * n));
                const float c0 = (float)p.x, c1 = (float)p.y;
                float prev = __shfl_up_sync(FULL, c1, 1);
                if (lane == 0) prev = carry;
                carry = __shfl_sync(FULL, c1, 31);
                const float y0 = preemph == 0.f ? c0 : __fsub_rn(c0, __fmul_rn(preemph, prev));
                const float y1 = preemph == 0.f ? c1 : __fsub_rn(c1, __fmul_rn(preemph, c0));
                a[padE(n)] = make_float2(sm.win[2 * n] * y0, sm.win[2 * n + 1] * y1);
            }
        } else {
            for (int n = lane; n < M; n += 32)
                a[padE(n)] = make_float2(sm.win[2 * n] * emph(base + 2 * n), sm.win[2 * n + 1] * emph(base + 2 * n + 1));
        }
        __syncwarp();
        // Stockham autosort FFT of M points: radix-4 passes (sub-transform size Ns = 1, 4, 16, ...), one radix-2 pass at
        // the end when log2 M is odd.  Pass: v[r] = in[j + r M/R] W^(r (j mod Ns) M / (Ns R)); out[expand(j) + r Ns] = DFT_R(v)[r]
        float2* in = a;
        float2* out = b;
        int Ns = 1;
#pragma unroll 1
        for (; Ns * 4 <= M; Ns *= 4) {
            const int tw_step = M / (Ns * 4);
            for (int j = lane; j < M / 4; j += 32) {
                const int k = j & (Ns - 1);
                float2 v0 = in[padE(j)], v1 = in[padE(j + M / 4)], v2 = in[padE(j + M / 2)], v3 = in[padE(j + 3 * M / 4)];
                if (Ns > 1) {
                    v1 = cmulE(v1, sm.wm[k * tw_step]);
                    v2 = cmulE(v2, sm.wm[2 * k * tw_step]);
                    v3 = cmulE(v3, sm.wm[3 * k * tw_step]);
                }
                const float2 s0_ = make_float2(v0.x + v2.x, v0.y + v2.y), s1_ = make_float2(v0.x - v2.x, v0.y - v2.y);
                const float2 s2_ = make_float2(v1.x + v3.x, v1.y + v3.y), s3_ = make_float2(v1.x - v3.x, v1.y - v3.y);
                const int j0 = ((j - k) << 2) + k;
                out[padE(j0)] = make_float2(s0_.x + s2_.x, s0_.y + s2_.y);
                out[padE(j0 + Ns)] = make_float2(s1_.x + s3_.y, s1_.y - s3_.x);          // s1 - i s3
                out[padE(j0 + 2 * Ns)] = make_float2(s0_.x - s2_.x, s0_.y - s2_.y);
                out[padE(j0 + 3 * Ns)] = make_float2(s1_.x - s3_.y, s1_.y + s3_.x);      // s1 + i s3
            }
            __syncwarp();
            float2* tmp = in; in = out; out = tmp;
        }
        if (Ns < M) {                                   // one radix-2 pass left (Ns == M / 2)
            for (int j = lane; j < M / 2; j += 32) {
                const float2 v0 = in[padE(j)], v1 = cmulE(in[padE(j + M / 2)], sm.wm[j]);          // k = j, step = M / (2 Ns) = 1
                out[padE(j)] = make_float2(v0.x + v1.x, v0.y + v1.y);
                out[padE(j + Ns)] = make_float2(v0.x - v1.x, v0.y - v1.y);
            }
            __syncwarp();
            float2* tmp = in; in = out; out = tmp;
        }
        // real-input post-pass: X[k] = (E + W_N^k O) / 2 with E = Z[k] + conj Z[M-k], O = -i (Z[k] - conj Z[M-k]);
        // power spectrum into the other buffer (M + 1 floats, unpadded)
        float* pw = reinterpret_cast<float*>(out);
        for (int k = lane; k <= M; k += 32) {
            const float2 A = in[padE(k & (M - 1))], B = in[padE((M - k) & (M - 1))];
            const float2 e = make_float2(A.x + B.x, A.y - B.y);
            const float2 o = make_float2(A.y + B.y, B.x - A.x);
            const float2 wo = cmulE(sm.wn[k], o);
            const float xr = e.x + wo.x, xi = e.y + wo.y;
            pw[k] = 0.25f * (xr * xr + xi * xi);
        }
        __syncwarp();
        // triangular filters from the lane table (a zero weight may sit on a bin past the filter's end: finite, times 0)
        float accA = 0.f, accB = 0.f;
        for (int it = 0; it < na; ++it) accA = fmaf(s_w[it * 32 + lane], pw[min(binA + it, M)], accA);
        for (int it = 0; it < nb; ++it) accB = fmaf(s_w[(na + it) * 32 + lane], pw[min(binB + lb * it, M)], accB);
        for (int o = 1; o < lb; o <<= 1) accB += __shfl_xor_sync(FULL, accB, o);
        float* mo = mel_out + (f0 + t) * n_mels;
        if (lane < nA) { mo[lane] = accA; vmax = fmaxf(vmax, accA); }
        if (lane % lb == 0 && lane / lb < nB) { mo[32 + lane / lb] = accB; vmax = fmaxf(vmax, accB); }
        __syncwarp();
    }
    if (utt_max) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(FULL, vmax, o));
        if (lane == 0) atomicMax(reinterpret_cast<int*>(utt_max + u), __float_as_int(vmax));        // mel >= 0
    }
}

// ------------------------------------------------------------------------------------------------------------------
// mel_ex512_kernel: the 512-point case (the "spec" parameter set) in the real-input-first form of mfcc_mel_r_kernel
// (mfcc.cu): 512 = 32 x 16, n = 16 n1 + n2, k = k1 + 32 k2.
//   step 1, thread (frame, n2):  T[k1][n2] = W_512^(n2 k1) * sum_n1 w[n] y[n] W_32^(n1 k1) for k1 = 0..16 -- a 32-point real
//            DFT in registers: 16-point complex FFT of z[m] = (y[2m], y[2m+1]) and the in-thread real-input post-pass
//            (compile-time W_32 twiddles; its factor 1/2 rides in the W_512 twiddles);
//   step 2, thread (frame, k1):  one 16-point FFT of the row = the 16 bins k1 + 32 k2 (mirrored above 256).
// Warp-private batches of 8 frames, 17 rows x 8 frames x (128 + 16) bytes per warp; the power planes ([plane of 4 frames]
// [bin][frame]) take the place of the rows 9..16, so step 2 runs the rows 13..16 first (powers kept in registers), then
// 9..12, 5..8, 1..4 and 0.  The filterbank runs over the 8 frames at once from the lane table of mel_ex_kernel.
// Pre-emphasis needs the sample before every sample: a second set of loads (the window then comes from shared memory
// instead of registers).  Samples are requested without tests from a base clamped into the utterance as soon as the
// previous item has consumed its own; frames reaching outside the utterance are fetched again with bounds tests.
__device__ float2 g_w512[256];                   // W_512^j, j < 256 (uploaded once per device, see ensure_w512)

namespace r512 {
constexpr int kWarps = 4;
constexpr int kBatch = 8;
constexpr int kN = 512;
constexpr int kBinsN = kN / 2 + 1;
constexpr int kRows = 17;
constexpr int kFrameB = 144;
constexpr int kRowB = kBatch * kFrameB;          // 1152
constexpr int kAreaB = kRows * kRowB;            // 19584 bytes per warp
constexpr int kPlaneB = 4160;                    // 260 bins x 16 bytes; 16 banks mod 32
constexpr int kPlane0 = 9 * kRowB;
constexpr int kWinPitch = 34;                    // floats per n2 row of the permuted window: the 8-byte loads of 16 lanes hit 32 banks
static_assert(kPlane0 + 2 * kPlaneB <= kAreaB && kBinsN * 16 <= kPlaneB && (kPlaneB / 4) % 32 == 16, "power planes");

struct __align__(16) Smem {
    unsigned char area[kWarps][kAreaB];
    float win[16 * kWinPitch];                   // win[n2 * kWinPitch + n1] = window[16 n1 + n2]
    int binA[32], binB[32];
    int na, nb, lb;
};
}  // namespace r512

template <typename SampleT, bool PREEMPH>
__global__ void __launch_bounds__(r512::kWarps * 32, 2)
mel_ex512_kernel(const SampleT* __restrict__ pcm, const int64_t* __restrict__ pcm_off, const int64_t* __restrict__ frm_off,
                 const float* __restrict__ window, int hop, float preemph,
                 const int32_t* __restrict__ mel_start, const int32_t* __restrict__ mel_len, const float* __restrict__ mel_w,
                 int mel_pitch, int n_mels, int chunk, float* __restrict__ mel_out, float* __restrict__ utt_max) {
    using namespace r512;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int kThreads = kWarps * 32;
    extern __shared__ __align__(16) unsigned char smem_raw_e[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw_e);
    float* s_w = reinterpret_cast<float*>(smem_raw_e + sizeof(Smem));     // [(na + nb) * 32] lane table of filter weights
    const int u = blockIdx.x;
    const int64_t f0 = frm_off[u];
    const int T = (int)(frm_off[u + 1] - f0);
    const int t_begin = blockIdx.y * chunk;
    if (t_begin >= T) return;
    const int t_end = min(T, t_begin + chunk);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < kN; i += kThreads) sm.win[(i & 15) * kWinPitch + (i >> 4)] = window[i];
    for (int i = tid; i < kWarps * kAreaB / 4; i += kThreads) reinterpret_cast<float*>(&sm.area[0][0])[i] = 0.f;
    // filterbank as a lane table, as mel_ex_kernel builds it
    const int nA = min(n_mels, 32), nB = n_mels - nA;
    if (warp == 0) {
        int la = lane < nA ? mel_len[lane] : 0;
        int lbn = lane < nB ? mel_len[32 + lane] : 0;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) { la = max(la, __shfl_xor_sync(FULL, la, o)); lbn = max(lbn, __shfl_xor_sync(FULL, lbn, o)); }
        int lb = 32;
        while (lb > 1 && lb * nB > 32) lb >>= 1;
        if (lane == 0) { sm.na = la; sm.lb = lb; sm.nb = nB ? (lbn + lb - 1) / lb : 0; }
    }
    __syncthreads();
    const int na = sm.na, nb = sm.nb, lb = sm.lb;
    for (int i = tid; i < na * 32; i += kThreads) {
        const int it = i >> 5, l = i & 31;
        s_w[i] = (l < nA && it < mel_len[l]) ? mel_w[(size_t)l * mel_pitch + it] : 0.f;
    }
    for (int i = tid; i < nb * 32; i += kThreads) {
        const int it = i >> 5, l = i & 31, q = l / lb, j = l % lb, m = 32 + q, e = j + lb * it;
        s_w[na * 32 + i] = (q < nB && e < mel_len[m]) ? mel_w[(size_t)m * mel_pitch + e] : 0.f;
    }
    if (tid < 32) {
        sm.binA[tid] = tid < nA ? mel_start[tid] : 0;
        const int q = tid / lb;
        sm.binB[tid] = q < nB ? mel_start[32 + q] + tid % lb : 0;
    }
    __syncthreads();
    const int binA = sm.binA[lane], binB = sm.binB[lane];

    // step 1 constants, thread = (frame parity fl, n2)
    const int n2 = lane & 15, fl = lane >> 4;
    float2 hw[PREEMPH ? 1 : 16];                       // window taps (w[16 (2m) + n2], w[16 (2m + 1) + n2]); with pre-emphasis: shared memory
    if (!PREEMPH) {
#pragma unroll
        for (int m = 0; m < 16; ++m) hw[m] = make_float2(sm.win[n2 * kWinPitch + 2 * m], sm.win[n2 * kWinPitch + 2 * m + 1]);
    }
    float2 tw[17];                                      // W_512^(n2 k1), halved for k1 = 1..15 (the real-input post-pass)
#pragma unroll
    for (int k1 = 1; k1 <= 16; ++k1) {
        const float2 wv = g_w512[n2 * k1];              // n2 k1 <= 240
        const float h = k1 < 16 ? 0.5f : 1.0f;
        tw[k1] = make_float2(h * wv.x, h * wv.y);
    }
    const int f8 = lane & 7, kq = lane >> 3;

    const int64_t s0 = pcm_off[u];
    const int ns = (int)(pcm_off[u + 1] - s0);
    const SampleT* __restrict__ x = pcm + s0;
    unsigned char* area = sm.area[warp];
    float vmax = 0.f;

    constexpr int kLo = PREEMPH ? 1 : 0;               // an interior frame also owns the sample before its first
    const int base_max = ns - kN;                       // >= kLo (the caller checks the shortest utterance)
    auto edge = [&](int t) { const int base = hop * t - kN / 2; return base < kLo || base > base_max; };
    auto fetch_interior = [&](int t, float* r, float* p) {
        const SampleT* __restrict__ xb = x + min(max(hop * t - kN / 2, kLo), base_max) + n2;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) r[n1] = to_f32(__ldg(xb + 16 * n1));
        if (PREEMPH) {
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) p[n1] = to_f32(__ldg(xb + 16 * n1 - 1));
        }
    };
    auto fetch_edge = [&](int t, float* r, float* p) {   // zero outside the utterance; the first sample is its own predecessor
        const int base = hop * t - kN / 2 + n2;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const int i = base + 16 * n1;
            const bool in = i >= 0 && i < ns;
            r[n1] = in ? to_f32(__ldg(x + i)) : 0.f;
            if (PREEMPH) p[n1] = in ? to_f32(__ldg(x + (i > 0 ? i - 1 : 0))) : 0.f;
        }
    };
    float r[32], p[PREEMPH ? 32 : 1];
    fetch_interior(t_begin + warp * kBatch + fl, r, p);

    for (int tb = t_begin + warp * kBatch; tb < t_end; tb += kWarps * kBatch) {
        // ---------------- step 1
#pragma unroll 1
        for (int ps = 0; ps < kBatch / 2; ++ps) {
            const int fb = 2 * ps + fl, t = tb + fb;
            if (edge(t) && t < t_end) fetch_edge(t, r, p);
            float2 z[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                float y0 = r[2 * m], y1 = r[2 * m + 1];
                if (PREEMPH) {                         // never contracted: float32 and int16 PCM agree bit for bit
                    y0 = __fsub_rn(y0, __fmul_rn(preemph, p[2 * m]));
                    y1 = __fsub_rn(y1, __fmul_rn(preemph, p[2 * m + 1]));
                }
                float2 w;
                if (PREEMPH) w = *reinterpret_cast<const float2*>(&sm.win[n2 * kWinPitch + 2 * m]);
                else w = hw[PREEMPH ? 0 : m];
                z[m] = emul(make_float2(y0, y1), w);
            }
            fetch_interior((ps < kBatch / 2 - 1) ? t + 2 : tb + kWarps * kBatch + fl, r, p);
            if (t < t_end) {
                fft16(z);
                // real-input post-pass: 2 Y[k] = e + w o, 2 Y[16 - k] = conj(e - w o), e = Z[k] + conj Z[16 - k],
                // o = -i (Z[k] - conj Z[16 - k]), w = W_32^k
                float2 Y[17];
                Y[0] = make_float2(z[0].x + z[0].y, 0.f);
                Y[16] = make_float2(z[0].x - z[0].y, 0.f);
                Y[8] = make_float2(2.f * z[8].x, -2.f * z[8].y);
                {
                    constexpr float C32[8] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                                              0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f};
                    constexpr float S32[8] = {0.f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f,
                                              0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f};
#pragma unroll
                    for (int k = 1; k < 8; ++k) {
                        const float2 A = z[k], B = z[16 - k];
                        const float2 e = make_float2(A.x + B.x, A.y - B.y);
                        const float2 o = make_float2(A.y + B.y, B.x - A.x);
                        const float2 wo = cmul(make_float2(C32[k], -S32[k]), o);
                        Y[k] = cadd(e, wo);
                        const float2 d = csub(e, wo);
                        Y[16 - k] = make_float2(d.x, -d.y);
                    }
                }
                float2* slot = reinterpret_cast<float2*>(area + fb * kFrameB) + n2;
                slot[0] = Y[0];
#pragma unroll
                for (int k1 = 1; k1 < 16; ++k1) slot[k1 * (kRowB / 8)] = cmul(Y[k1], tw[k1]);
                slot[16 * (kRowB / 8)] = make_float2(Y[16].x * tw[16].x, Y[16].x * tw[16].y);
            }
        }
        __syncwarp();
        // ---------------- step 2
        const bool fvalid = tb + f8 < t_end;
        float* plane = reinterpret_cast<float*>(area + kPlane0 + (f8 >> 2) * kPlaneB) + (f8 & 3);
        auto load_row = [&](int k1, float2* zz) {
            const float4* src = reinterpret_cast<const float4*>(area + k1 * kRowB + f8 * kFrameB);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 q = src[c];
                zz[2 * c] = make_float2(q.x, q.y);
                zz[2 * c + 1] = make_float2(q.z, q.w);
            }
        };
        auto power = [&](const float2* zz, float* pw) {
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) pw[k2] = fmaf(zz[k2].x, zz[k2].x, zz[k2].y * zz[k2].y);
        };
        auto store_pw = [&](int k1, const float* pw) {
            float* lo = plane + 4 * k1;                    // bin k1 + 32 k2, k2 = 0..7
            float* hi = plane + 4 * (kN - k1);             // bin 512 - k1 - 32 k2, k2 = 8..15
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) lo[128 * k2] = pw[k2];
#pragma unroll
            for (int k2 = 8; k2 < 16; ++k2) hi[-128 * k2] = pw[k2];
        };
        {
            float pc[16], pd[16];
            float2 za[16], zb[16];
            if (fvalid) { load_row(13 + kq, za); fft16(za); power(za, pc); }
            if (fvalid) load_row(9 + kq, zb);
            __syncwarp();                                   // rows 9..16 are in registers: their area takes the power planes
            if (fvalid) {
                store_pw(13 + kq, pc);
                fft16(zb); power(zb, pd); store_pw(9 + kq, pd);
                load_row(5 + kq, za); fft16(za); power(za, pc); store_pw(5 + kq, pc);
                load_row(1 + kq, zb); fft16(zb); power(zb, pd); store_pw(1 + kq, pd);
                if (kq == 0) { load_row(0, za); fft16(za); power(za, pc); store_pw(0, pc); }
            }
        }
        __syncwarp();
        // ---------------- mel filterbank, 8 frames at once (run-time trip counts: any number of filters up to 64)
        {
            const int nf = min(kBatch, t_end - tb);
            const unsigned char* planes = area + kPlane0;
            float acc[8];
#pragma unroll
            for (int f = 0; f < 8; ++f) acc[f] = 0.f;
            for (int it = 0; it < na; ++it) {
                const float w = s_w[it * 32 + lane];
                const unsigned char* pb_ = planes + min(binA + it, kN / 2) * 16;
                const float4 a4 = *reinterpret_cast<const float4*>(pb_);
                const float4 b4 = *reinterpret_cast<const float4*>(pb_ + kPlaneB);
                acc[0] = fmaf(w, a4.x, acc[0]); acc[1] = fmaf(w, a4.y, acc[1]); acc[2] = fmaf(w, a4.z, acc[2]); acc[3] = fmaf(w, a4.w, acc[3]);
                acc[4] = fmaf(w, b4.x, acc[4]); acc[5] = fmaf(w, b4.y, acc[5]); acc[6] = fmaf(w, b4.z, acc[6]); acc[7] = fmaf(w, b4.w, acc[7]);
            }
            if (lane < nA) {
                float* mo = mel_out + (f0 + tb) * n_mels + lane;
#pragma unroll
                for (int f = 0; f < 8; ++f) {
                    if (f < nf) { *mo = acc[f]; vmax = fmaxf(vmax, acc[f]); }
                    mo += n_mels;
                }
            }
            if (nb > 0) {
#pragma unroll
                for (int f = 0; f < 8; ++f) acc[f] = 0.f;
                for (int it = 0; it < nb; ++it) {
                    const float w = s_w[(na + it) * 32 + lane];
                    const unsigned char* pb_ = planes + min(binB + lb * it, kN / 2) * 16;
                    const float4 a4 = *reinterpret_cast<const float4*>(pb_);
                    const float4 b4 = *reinterpret_cast<const float4*>(pb_ + kPlaneB);
                    acc[0] = fmaf(w, a4.x, acc[0]); acc[1] = fmaf(w, a4.y, acc[1]); acc[2] = fmaf(w, a4.z, acc[2]); acc[3] = fmaf(w, a4.w, acc[3]);
                    acc[4] = fmaf(w, b4.x, acc[4]); acc[5] = fmaf(w, b4.y, acc[5]); acc[6] = fmaf(w, b4.z, acc[6]); acc[7] = fmaf(w, b4.w, acc[7]);
                }
                for (int o = 1; o < lb; o <<= 1) {
#pragma unroll
                    for (int f = 0; f < 8; ++f) acc[f] += __shfl_xor_sync(FULL, acc[f], o);
                }
                if (lane % lb == 0 && lane / lb < nB) {
                    float* mb = mel_out + (f0 + tb) * n_mels + 32 + lane / lb;
#pragma unroll
                    for (int f = 0; f < 8; ++f) {
                        if (f < nf) { *mb = acc[f]; vmax = fmaxf(vmax, acc[f]); }
                        mb += n_mels;
                    }
                }
            }
        }
        __syncwarp();
    }
    if (utt_max) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(FULL, vmax, o));
        if (lane == 0) atomicMax(reinterpret_cast<int*>(utt_max + u), __float_as_int(vmax));        // mel >= 0
    }
}

// thread per frame: log, DCT-II (coefficient table in shared memory).  NM / NC > 0: compile-time sizes (the mel row and
// the loops live in registers); 0: run-time sizes (local-memory row).
template <int NM, int NC>
__global__ void __launch_bounds__(128)
ceps_ex_kernel(const float* __restrict__ mel, const float* __restrict__ utt_max, const int64_t* __restrict__ frm_off, int n_utt,
               int64_t total_frames, const float* __restrict__ dct, int n_mels_rt, int n_ceps_rt, int log_mode, float* __restrict__ ceps) {
    const int n_mels = NM > 0 ? NM : n_mels_rt, n_ceps = NC > 0 ? NC : n_ceps_rt;
    __shared__ float s_dct[kMaxCepsE * kMaxMelsE];
    for (int i = threadIdx.x; i < n_ceps * n_mels; i += blockDim.x) s_dct[i] = dct[i];
    __syncthreads();
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= total_frames) return;
    // 10 log10(x) = kDbPerLog2E log2(x), ln(x) = kLnPerLog2E log2(x): MUFU.LG2 (2 ulp) instead of the 20-instruction libm calls,
    // as mfcc_ceps_kernel does
    constexpr float kDbPerLog2E = 3.0102999566398120f, kLnPerLog2E = 0.69314718055994531f;
    float ref_db = 0.f;
    if (log_mode == LOE_LOG_DB) {
        int lo = 0, hi = n_utt;                    // utterance of this frame: last u with frm_off[u] <= f
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (frm_off[mid] <= f) lo = mid; else hi = mid; }
        ref_db = kDbPerLog2E * __log2f(fmaxf(1e-10f, utt_max[lo]));
    }
    float lm[NM > 0 ? NM : kMaxMelsE];
    const float* row = mel + f * n_mels;
    if (NM > 0) {
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            const float e = fmaxf(1e-10f, __ldg(row + m));
            lm[m] = log_mode == LOE_LOG_DB ? fmaxf(fmaf(kDbPerLog2E, __log2f(e), -ref_db), -80.0f) : kLnPerLog2E * __log2f(e);
        }
    } else {
#pragma unroll 1
        for (int m = 0; m < n_mels; ++m) {
            const float e = fmaxf(1e-10f, __ldg(row + m));
            lm[m] = log_mode == LOE_LOG_DB ? fmaxf(fmaf(kDbPerLog2E, __log2f(e), -ref_db), -80.0f) : kLnPerLog2E * __log2f(e);
        }
    }
    float* o = ceps + f * n_ceps;
    if (NM > 0) {
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            float acc = 0.f;
#pragma unroll
            for (int m = 0; m < NM; ++m) acc = fmaf(s_dct[k * NM + m], lm[m], acc);
            o[k] = acc;
        }
    } else {
        for (int k = 0; k < n_ceps; ++k) {
            float acc = 0.f;
            for (int m = 0; m < n_mels; ++m) acc = fmaf(s_dct[k * n_mels + m], lm[m], acc);
            o[k] = acc;
        }
    }
}

// CTA per utterance: mean and 1 / (std + 1e-8) of every coefficient over the utterance's frames.  Thread t sums the
// frames t, t + 128, ... in float64, the partials are combined in a fixed tree: bitwise reproducible.
__global__ void __launch_bounds__(128)
cmn_ex_kernel(const float* __restrict__ ceps, const int64_t* __restrict__ frm_off, int n_ceps, float* __restrict__ stat) {
    __shared__ double s_sum[128], s_sq[128];
    const int u = blockIdx.x;
    const int64_t f0 = frm_off[u];
    const int T = (int)(frm_off[u + 1] - f0);
    for (int k = 0; k < n_ceps; ++k) {
        double a = 0.0, q = 0.0;
        for (int t = threadIdx.x; t < T; t += 128) { const double v = ceps[(f0 + t) * n_ceps + k]; a += v; q += v * v; }
        s_sum[threadIdx.x] = a; s_sq[threadIdx.x] = q;
        __syncthreads();
        for (int o = 64; o >= 1; o >>= 1) {
            if (threadIdx.x < o) { s_sum[threadIdx.x] += s_sum[threadIdx.x + o]; s_sq[threadIdx.x] += s_sq[threadIdx.x + o]; }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const double mean = s_sum[0] / T;
            const double var = fmax(0.0, s_sq[0] / T - mean * mean);
            stat[(size_t)u * 2 * n_ceps + k] = (float)mean;
            stat[(size_t)u * 2 * n_ceps + n_ceps + k] = (float)(1.0 / (sqrt(var) + 1e-8));
        }
        __syncthreads();
    }
}

// thread per frame: normalised static block, Savitzky-Golay delta / delta-delta of width 9 over the raw cepstra (edge
// frames take the value of the nearest full window: scipy savgol_filter mode="interp" with polyorder == deriv)
__global__ void __launch_bounds__(128)
feat_ex_kernel(const float* __restrict__ ceps, const float* __restrict__ stat, const int64_t* __restrict__ frm_off, int n_utt,
               int64_t total_frames, int n_ceps, int norm_mode, float* __restrict__ feat) {
    // a CTA's 128 frames are 128 consecutive rows of the feature matrix whatever utterances they belong to: the rows are
    // put together in shared memory (odd pitch: row-per-lane stores hit 32 banks) and written out as one contiguous run
    __shared__ float s_out[128 * (3 * kMaxCepsE + 1)];
    const int width = 3 * n_ceps, pitch = width | 1;
    const int64_t fb = (int64_t)blockIdx.x * blockDim.x;
    const int64_t f = fb + threadIdx.x;
    if (f < total_frames) {
        int lo = 0, hi = n_utt;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (frm_off[mid] <= f) lo = mid; else hi = mid; }
        const int64_t f0 = frm_off[lo];
        const int T = (int)(frm_off[lo + 1] - f0);
        const int t = (int)(f - f0);
        const float* c = ceps + f * n_ceps;
        float* o = s_out + threadIdx.x * pitch;
        if (norm_mode == LOE_NORM_FRAME) {
            float mean = 0.f;
            for (int k = 0; k < n_ceps; ++k) mean += c[k];
            mean /= (float)n_ceps;
            float var = 0.f;
            for (int k = 0; k < n_ceps; ++k) { const float d = c[k] - mean; var = fmaf(d, d, var); }
            const float inv = 1.0f / (sqrtf(var / (float)n_ceps) + 1e-8f);
            for (int k = 0; k < n_ceps; ++k) o[k] = (c[k] - mean) * inv;
        } else if (norm_mode == LOE_NORM_CMN || norm_mode == LOE_NORM_CMVN) {
            const float* st = stat + (size_t)lo * 2 * n_ceps;
            for (int k = 0; k < n_ceps; ++k) {
                const float d = c[k] - st[k];
                o[k] = norm_mode == LOE_NORM_CMVN ? d * st[n_ceps + k] : d;
            }
        } else {
            for (int k = 0; k < n_ceps; ++k) o[k] = c[k];
        }
        const int cc = min(max(t, 4), T - 5);
        float d1[kMaxCepsE], d2[kMaxCepsE];
#pragma unroll
        for (int k = 0; k < kMaxCepsE; ++k) { d1[k] = 0.f; d2[k] = 0.f; }
#pragma unroll
        for (int q = -4; q <= 4; ++q) {                   // neighbour-major: every row of 13 coefficients is read as a run
            const float* nrow = ceps + (f0 + cc + q) * n_ceps;
            const float w1 = (float)q * (1.0f / 60.0f), w2 = (float)(3 * q * q - 20) * (1.0f / 462.0f);
#pragma unroll
            for (int k = 0; k < kMaxCepsE; ++k)
                if (k < n_ceps) {
                    const float cv = __ldg(nrow + k);
                    d1[k] = fmaf(w1, cv, d1[k]);
                    d2[k] = fmaf(w2, cv, d2[k]);
                }
        }
#pragma unroll
        for (int k = 0; k < kMaxCepsE; ++k)
            if (k < n_ceps) { o[n_ceps + k] = d1[k]; o[2 * n_ceps + k] = d2[k]; }
    }
    __syncthreads();
    const int rows = (int)min((int64_t)blockDim.x, total_frames - fb);
    float* __restrict__ dst = feat + fb * width;
    for (int i = threadIdx.x; i < rows * width; i += blockDim.x) {
        const int rr = i / width;
        dst[i] = s_out[rr * pitch + (i - rr * width)];
    }
}

template <typename SampleT, int LOG2N>
static int launch_mel_ex(const void* pcm_dev, const int64_t* pcm_off_dev, const int64_t* frm_off_dev, int n_utt, int max_frames,
                         int chunk, const loe_mfcc_config* cfg, const float* window_dev, const int32_t* mel_start_dev,
                         const int32_t* mel_len_dev, const float* mel_w_dev, int mel_pitch, float* mel_ws_dev, float* utt_max_dev,
                         cudaStream_t s) {
    using S = SmemE<LOG2N>;
    const size_t smem = sizeof(S) + sizeof(float) * 64 * (size_t)mel_pitch;          // + lane table: (na + nb) * 32 <= 2 * pitch * 32
    LOE_CUDA(cudaFuncSetAttribute(mel_ex_kernel<SampleT, LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)n_utt, (unsigned)((max_frames + chunk - 1) / chunk));
    mel_ex_kernel<SampleT, LOG2N><<<grid, kWarpsE * 32, smem, s>>>(
        (const SampleT*)pcm_dev, pcm_off_dev, frm_off_dev, window_dev, cfg->hop, cfg->preemph, mel_start_dev, mel_len_dev, mel_w_dev,
        mel_pitch, cfg->n_mels, chunk, mel_ws_dev, cfg->log_mode == LOE_LOG_DB ? utt_max_dev : nullptr);
    LOE_LAUNCH_CHECK("mel_ex_kernel");
    return LOE_OK;
}

// one-time upload of W_512^j per device (double-precision host values; the copy goes through the legacy stream from
// pageable memory while the kernel runs on the caller's stream: the device is synchronised once, host threads serialised)
static int ensure_w512() {
    static std::atomic<bool> ready[64];
    static std::mutex mtx;
    int dev = 0;
    LOE_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && ready[dev].load(std::memory_order_acquire)) return LOE_OK;
    std::lock_guard<std::mutex> lock(mtx);
    if (dev < 64 && ready[dev].load(std::memory_order_acquire)) return LOE_OK;
    static float2 h[256];
    for (int j = 0; j < 256; ++j) {
        const double a = -2.0 * 3.14159265358979323846 * j / 512.0;
        h[j] = make_float2((float)cos(a), (float)sin(a));
    }
    LOE_CUDA(cudaMemcpyToSymbol(g_w512, h, sizeof(h)));
    LOE_CUDA(cudaDeviceSynchronize());
    if (dev < 64) ready[dev].store(true, std::memory_order_release);
    return LOE_OK;
}

template <typename SampleT, bool PREEMPH>
static int launch_mel_ex512(const void* pcm_dev, const int64_t* pcm_off_dev, const int64_t* frm_off_dev, int n_utt, int max_frames,
                            int chunk, const loe_mfcc_config* cfg, const float* window_dev, const int32_t* mel_start_dev,
                            const int32_t* mel_len_dev, const float* mel_w_dev, int mel_pitch, float* mel_ws_dev, float* utt_max_dev,
                            cudaStream_t s) {
    const size_t smem = sizeof(r512::Smem) + sizeof(float) * 64 * (size_t)mel_pitch;
    const int st = ensure_w512();
    if (st != LOE_OK) return st;
    LOE_CUDA(cudaFuncSetAttribute(mel_ex512_kernel<SampleT, PREEMPH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)n_utt, (unsigned)((max_frames + chunk - 1) / chunk));
    mel_ex512_kernel<SampleT, PREEMPH><<<grid, r512::kWarps * 32, smem, s>>>(
        (const SampleT*)pcm_dev, pcm_off_dev, frm_off_dev, window_dev, cfg->hop, cfg->preemph, mel_start_dev, mel_len_dev, mel_w_dev,
        mel_pitch, cfg->n_mels, chunk, mel_ws_dev, cfg->log_mode == LOE_LOG_DB ? utt_max_dev : nullptr);
    LOE_LAUNCH_CHECK("mel_ex512_kernel");
    return LOE_OK;
}

}  // namespace loe

extern "C" int loe_mfcc_ex_dev(const void* pcm_dev, int pcm_format, const int64_t* pcm_off_dev, const int64_t* frm_off_dev,
                               int n_utt, int64_t total_frames, int max_frames, int min_frames, const loe_mfcc_config* cfg,
                               const float* window_dev, const int32_t* mel_start_dev, const int32_t* mel_len_dev,
                               const float* mel_w_dev, int mel_pitch, const float* dct_dev, float* mel_ws_dev, float* ceps_ws_dev,
                               float* utt_stat_dev, float* feat_dev, void* stream) {
    using namespace loe;
    if (n_utt <= 0 || total_frames <= 0) return LOE_OK;
    if (!cfg) { set_error("cfg is NULL"); return LOE_ERR_VALUE; }
    if (min_frames < 9) {
        set_error("MFCC needs at least 9 frames per utterance for the width-9 delta filter (got %d)", min_frames);
        return LOE_ERR_VALUE;
    }
    int log2n = 0;
    while ((1 << log2n) < cfg->n_fft) ++log2n;
    if ((1 << log2n) != cfg->n_fft || log2n < 6 || log2n > 10) {
        set_error("n_fft = %d: the parameterised front end takes a power of two in 64..1024 (320 is served by loe_mfcc_dev)", cfg->n_fft);
        return LOE_ERR_UNSUPPORTED;
    }
    if (cfg->hop <= 0 || cfg->n_mels <= 0 || cfg->n_mels > kMaxMelsE || cfg->n_ceps <= 0 || cfg->n_ceps > kMaxCepsE ||
        cfg->n_ceps > cfg->n_mels || mel_pitch <= 0 || mel_pitch > 256) {
        set_error("bad front-end sizes (hop %d, n_mels %d <= %d, n_ceps %d <= %d)", cfg->hop, cfg->n_mels, kMaxMelsE, cfg->n_ceps, kMaxCepsE);
        return LOE_ERR_VALUE;
    }
    if (cfg->log_mode != LOE_LOG_DB && cfg->log_mode != LOE_LOG_LN) { set_error("unknown log_mode %d", cfg->log_mode); return LOE_ERR_VALUE; }
    if (cfg->norm_mode < LOE_NORM_NONE || cfg->norm_mode > LOE_NORM_CMVN) { set_error("unknown norm_mode %d", cfg->norm_mode); return LOE_ERR_VALUE; }
    if (pcm_format != LOE_PCM_F32 && pcm_format != LOE_PCM_S16) { set_error("unknown pcm_format %d", pcm_format); return LOE_ERR_VALUE; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (cfg->log_mode == LOE_LOG_DB) LOE_CUDA(cudaMemsetAsync(utt_stat_dev, 0, sizeof(float) * (size_t)n_utt, s));
    int sms = 0, dev = 0;
    LOE_CUDA(cudaGetDevice(&dev));
    LOE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // frames per CTA: the per-CTA table set-up (twiddles, window) is paid once per chunk; keep ~8 CTAs per SM in flight or queued
    int64_t chunk64 = total_frames / (8 * (int64_t)sms);
    chunk64 = ((chunk64 + kWarpsE - 1) / kWarpsE) * kWarpsE;
    const int chunk = (int)(chunk64 < 64 ? 64 : chunk64 > (1 << 20) ? (1 << 20) : chunk64);
    int st = LOE_OK;
    // 512-point FFT (the "spec" set): the real-input-first register kernel, as long as the shortest utterance holds a whole
    // frame plus the sample before it (the kernel's test-free loads are clamped into the utterance)
    if (log2n == 9 && (int64_t)(min_frames - 1) * cfg->hop >= 514) {
        const int round = r512::kWarps * r512::kBatch;
        const int chunk5 = ((chunk + round - 1) / round) * round;
        const bool pre = cfg->preemph != 0.f;
        if (pcm_format == LOE_PCM_F32)
            st = pre ? launch_mel_ex512<float, true>(pcm_dev, pcm_off_dev, frm_off_dev, n_utt, max_frames, chunk5, cfg, window_dev, mel_start_dev,
                                                     mel_len_dev, mel_w_dev, mel_pitch, mel_ws_dev, utt_stat_dev, s)
                     : launch_mel_ex512<float, false>(pcm_dev, pcm_off_dev, frm_off_dev, n_utt, max_frames, chunk5, cfg, window_dev, mel_start_dev,
                                                      mel_len_dev, mel_w_dev, mel_pitch, mel_ws_dev, utt_stat_dev, s);
        else
            st = pre ? launch_mel_ex512<short, true>(pcm_dev, pcm_off_dev, frm_off_dev, n_utt, max_frames, chunk5, cfg, window_dev, mel_start_dev,
                                                     mel_len_dev, mel_w_dev, mel_pitch, mel_ws_dev, utt_stat_dev, s)
                     : launch_mel_ex512<short, false>(pcm_dev, pcm_off_dev, frm_off_dev, n_utt, max_frames, chunk5, cfg, window_dev, mel_start_dev,
                                                      mel_len_dev, mel_w_dev, mel_pitch, mel_ws_dev, utt_stat_dev, s);
    } else {
#define LOE_EX(L)                                                                                                                    \
    case L:                                                                                                                          \
        st = pcm_format == LOE_PCM_F32                                                                                               \
                 ? launch_mel_ex<float, L>(pcm_dev, pcm_off_dev, frm_off_dev, n_utt, max_frames, chunk, cfg, window_dev, mel_start_dev, \
                                           mel_len_dev, mel_w_dev, mel_pitch, mel_ws_dev, utt_stat_dev, s)                            \
                 : launch_mel_ex<short, L>(pcm_dev, pcm_off_dev, frm_off_dev, n_utt, max_frames, chunk, cfg, window_dev, mel_start_dev, \
                                           mel_len_dev, mel_w_dev, mel_pitch, mel_ws_dev, utt_stat_dev, s);                           \
        break
    switch (log2n) { LOE_EX(6); LOE_EX(7); LOE_EX(8); LOE_EX(9); LOE_EX(10); }
#undef LOE_EX
    }
    if (st != LOE_OK) return st;
    const unsigned blocks = (unsigned)((total_frames + 127) / 128);
    if (cfg->n_mels == 40 && cfg->n_ceps == 13)
        ceps_ex_kernel<40, 13><<<blocks, 128, 0, s>>>(mel_ws_dev, utt_stat_dev, frm_off_dev, n_utt, total_frames, dct_dev, 40, 13,
                                                      cfg->log_mode, ceps_ws_dev);
    else
        ceps_ex_kernel<0, 0><<<blocks, 128, 0, s>>>(mel_ws_dev, utt_stat_dev, frm_off_dev, n_utt, total_frames, dct_dev, cfg->n_mels,
                                                    cfg->n_ceps, cfg->log_mode, ceps_ws_dev);
    LOE_LAUNCH_CHECK("ceps_ex_kernel");
    if (cfg->norm_mode == LOE_NORM_CMN || cfg->norm_mode == LOE_NORM_CMVN) {
        cmn_ex_kernel<<<(unsigned)n_utt, 128, 0, s>>>(ceps_ws_dev, frm_off_dev, cfg->n_ceps, utt_stat_dev);
        LOE_LAUNCH_CHECK("cmn_ex_kernel");
    }
    feat_ex_kernel<<<blocks, 128, 0, s>>>(ceps_ws_dev, utt_stat_dev, frm_off_dev, n_utt, total_frames, cfg->n_ceps, cfg->norm_mode, feat_dev);
    LOE_LAUNCH_CHECK("feat_ex_kernel");
    return LOE_OK;
}
