// MFCC front end for sm_100a.  Replaces mfcc.py:24-84 of the reference (librosa pipeline).
//
// Kernel A (mel):   PCM -> Hann window -> 320-point real DFT -> |.|^2 -> sparse slaney filterbank -> mel energies
//                   [frames, 40] + per-utterance maximum (atomicMax on the float bits).  Two forms, both with every
//                   sub-transform in the registers of one thread and ONE pass through shared memory:
//                   mfcc_mel_r_kernel (the 16 kHz filterbank table of every reference call site): real-input-first
//                   20 x 16 split, no real-input post-pass, filterbank over 8 frames at once;
//                   mfcc_mel_kernel (any other table): 160-point complex FFT split 10 x 16 + real-input post-pass.
// Kernel B (ceps):  mel -> dB relative to the utterance maximum, floor at -80 dB -> DCT-II
//                   (ortho) 13 ceps -> Savitzky-Golay delta / delta-delta (width 9, edge
//                   frames take the value of the nearest full window) -> per-frame
//                   normalisation of the static block -> features [frames, 39].
// The split is forced by power_to_db(ref=np.max): every frame needs the maximum over the whole
// utterance (mfcc.py:35).  Algorithmic HBM bytes per frame: 640 (PCM) + 156 (features).
#include "common.cuh"
#include "h16_stage.cuh"
#include "fft_regs.cuh"
#include <math.h>
#include <atomic>
#include <mutex>

namespace loe {

constexpr int kNfft = LOE_N_FFT;       // 320
constexpr int kHop = LOE_HOP;          // 160
constexpr int kHalf = kNfft / 2;       // 160 complex points
constexpr int kBins = LOE_N_BINS;      // 161
constexpr int kMels = LOE_N_MELS;      // 40
constexpr int kCeps = 13;
constexpr int kFeat = 39;

struct MfccTables {
    float hann[kNfft];                  // periodic Hann
    float w160_re[kHalf], w160_im[kHalf];   // W_160^j
    float w320_re[kBins], w320_im[kBins];   // W_320^k
    float dct[kCeps * kMels];           // ortho DCT-II rows
};

__device__ MfccTables g_mfcc_tables;
// kernel B reads these as FMA operands straight from the constant bank
__constant__ float c_dct[kCeps * kMels];        // ortho DCT-II rows
__constant__ float c_sg1[9], c_sg2[9];          // Savitzky-Golay first / second derivative taps, width 9
static std::atomic<bool> g_tables_ready[64];
static std::mutex g_tables_mutex;

// One-time upload of the constant tables per device.  The copies go through the legacy default stream from pageable
// memory while the kernels run on the caller's (possibly non-blocking) stream, which is not ordered after it: the
// device is synchronised once after the uploads, and the whole initialisation is serialised between host threads.
static int ensure_tables() {
    int dev = 0;
    LOE_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && g_tables_ready[dev].load(std::memory_order_acquire)) return LOE_OK;
    std::lock_guard<std::mutex> lock(g_tables_mutex);
    if (dev < 64 && g_tables_ready[dev].load(std::memory_order_acquire)) return LOE_OK;
    static MfccTables h;
    const double PI = 3.14159265358979323846;
    for (int n = 0; n < kNfft; ++n) h.hann[n] = (float)(0.5 - 0.5 * cos(2.0 * PI * n / kNfft));
    for (int j = 0; j < kHalf; ++j) {
        double a = -2.0 * PI * (double)j / kHalf;
        h.w160_re[j] = (float)cos(a);
        h.w160_im[j] = (float)sin(a);
    }
    for (int k = 0; k < kBins; ++k) {
        double a = -2.0 * PI * k / kNfft;
        h.w320_re[k] = (float)cos(a);
        h.w320_im[k] = (float)sin(a);
    }
    for (int k = 0; k < kCeps; ++k)
        for (int n = 0; n < kMels; ++n) {
            double v = sqrt(2.0 / kMels) * cos(PI * (2 * n + 1) * k / (2.0 * kMels));
            if (k == 0) v /= sqrt(2.0);
            h.dct[k * kMels + n] = (float)v;
        }
    LOE_CUDA(cudaMemcpyToSymbol(g_mfcc_tables, &h, sizeof(h)));
    LOE_CUDA(cudaMemcpyToSymbol(c_dct, h.dct, sizeof(h.dct)));
    float sg1[9], sg2[9];
    for (int q = -4; q <= 4; ++q) {
        sg1[q + 4] = (float)q * (1.0f / 60.0f);
        sg2[q + 4] = (float)(3 * q * q - 20) * (1.0f / 462.0f);
    }
    LOE_CUDA(cudaMemcpyToSymbol(c_sg1, sg1, sizeof(sg1)));
    LOE_CUDA(cudaMemcpyToSymbol(c_sg2, sg2, sizeof(sg2)));
    LOE_CUDA(cudaDeviceSynchronize());
    if (dev < 64) g_tables_ready[dev].store(true, std::memory_order_release);
    return LOE_OK;
}

// ------------------------------------------------------------------------------------------
// Kernel A
// ------------------------------------------------------------------------------------------
// 320-point real FFT = 160-point complex FFT of z[n] = x[2n] + i x[2n+1] plus a real-input post-pass.
// The 160-point transform is split 10 x 16 with every sub-transform held in the registers of ONE thread
// (no shuffles; the only exchange is one pass through shared memory):
//   n = 16 n1 + n2, k = k1 + 10 k2
//   step 1, thread (frame, n2):  B[k1][n2] = W_160^(n2 k1) * sum_n1 z[16 n1 + n2] W_10^(n1 k1)      (10-point, 2 x 5 prime-factor)
//   step 2, thread (frame, a):   Z[k1 + 10 k2] = sum_n2 B[k1][n2] W_16^(n2 k2) for k1 = a and k1 = 10 - a  (two radix-4 x 4 FFTs)
// The post-pass pairs bin k = a + 10 k2 with 160 - k = (10 - a) + 10 (15 - k2): both live in the same thread.
// a = 0 pairs with "k1 = 10", i.e. Z[10 + 10 k2]: step 1 stores an eleventh slot B[0][n2] W_16^n2, whose
// transform is the k1 = 0 output rotated by one; a = 5 pairs with itself.  Six threads per frame run the same
// code (a = 0 and a = 5 compute each of their pairs twice).
// A warp works on batches of 10 frames: step 1 in 5 passes of 2 frames, step 2 in 2 passes of 5 frames
// (30 lanes), then the mel filterbank one frame at a time with a filter per lane.
constexpr int kWarpsA = 4;
#ifndef LOE_MELR_MINB
#define LOE_MELR_MINB 3
#endif
#ifndef LOE_MEL_BATCH
#define LOE_MEL_BATCH 10
#endif
#ifndef LOE_MEL_MINB
#define LOE_MEL_MINB 3
#endif
constexpr int kBatchA = LOE_MEL_BATCH;                 // frames per warp batch
constexpr int kMinChunkA = 160;             // frames per CTA: at least this many
constexpr int kSlotPitch = 17;              // complex entries per slot (16 used): step 2's slot reads spread over the banks
constexpr int kSlots = 11;
constexpr int kFramePitch = 390;            // floats per frame: 11 x 17 x 2 = 374, padded to 6 mod 32 so that the
                                            // power-spectrum stores of 5 frames x 6 lanes hit 30 different banks
constexpr int kMelItMax = LOE_MEL_NA_MAX + LOE_MEL_NB_MAX;
static_assert(kSlots * kSlotPitch * 2 <= kFramePitch, "frame area too small");
static_assert(kBins + 3 + LOE_MEL_NA_MAX + 4 * LOE_MEL_NB_MAX <= kFramePitch, "power spectrum + table slack must fit the frame area");

struct __align__(16) SmemA {
    float area[kWarpsA][kBatchA * kFramePitch];     // per warp: step-1 slots, then the power spectra of the same frames
    float2 wpost[6 * kSlotPitch];                   // W_320^(a + 10 k2) at [a * 17 + k2]
    float mel_w[kMelItMax * 32];
};

template <typename SampleT> struct Pair;
template <> struct Pair<float> { using type = float2; };
template <> struct Pair<short> { using type = short2; };

// Mel filterbank as a lane-balanced table (host-built, see mfcc.py:mel_lane_tables):
//   round A, iterations [0, na):      lane l accumulates filter l          (filters 0..31)
//   round B, iterations [na, na+nb):  lanes 4q..4q+3 share filter 32 + q   (the 8 widest filters),
//                                     combined with two xor shuffles
// entry (it, lane) = weight mel_w[it*32+lane] applied to power bin mel_bin[it*32+lane] (0-weight padding).
template <typename SampleT, int NA, int NB>
__global__ void __launch_bounds__(kWarpsA * 32, LOE_MEL_MINB)
mfcc_mel_kernel(const SampleT* __restrict__ pcm, const int64_t* __restrict__ pcm_off,
                const int64_t* __restrict__ frm_off, const int32_t* __restrict__ mel_bin,
                const float* __restrict__ mel_w, int na_rt, int nb_rt, int chunk,
                float* __restrict__ mel_out, float* __restrict__ utt_max) {
    const int na = NA > 0 ? NA : na_rt;                  // NA, NB > 0: compile-time trip counts (loops unroll)
    const int nb = NA > 0 ? NB : nb_rt;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemA& sm = *reinterpret_cast<SmemA*>(smem_raw);
    const int u = blockIdx.x;
    const int64_t f0 = frm_off[u];
    const int T = (int)(frm_off[u + 1] - f0);
    const int t_begin = blockIdx.y * chunk;
    if (t_begin >= T) return;
    const int t_end = min(T, t_begin + chunk);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr unsigned FULL = 0xffffffffu;

    for (int i = tid; i < (na + nb) * 32; i += kWarpsA * 32) sm.mel_w[i] = mel_w[i];
    for (int i = tid; i < 6 * 16; i += kWarpsA * 32) {
        const int k = (i >> 4) + 10 * (i & 15);                  // a + 10 k2 <= 155
        sm.wpost[(i >> 4) * kSlotPitch + (i & 15)] = make_float2(g_mfcc_tables.w320_re[k], g_mfcc_tables.w320_im[k]);
    }
    // the frame areas start out finite: zero-weight filterbank entries may read slack words no step writes
    for (int i = tid; i < kWarpsA * kBatchA * kFramePitch; i += kWarpsA * 32) (&sm.area[0][0])[i] = 0.f;

    // per-lane constants of step 1 (thread = (frame parity, n2)): window taps and W_160^(n2 k1)
    const int n2 = lane & 15, fl = lane >> 4;
    float hw[20];
    float2 tw[10];
#pragma unroll
    for (int n1 = 0; n1 < 10; ++n1) {
        hw[2 * n1] = g_mfcc_tables.hann[32 * n1 + 2 * n2];
        hw[2 * n1 + 1] = g_mfcc_tables.hann[32 * n1 + 2 * n2 + 1];
        tw[n1] = make_float2(g_mfcc_tables.w160_re[n2 * n1], g_mfcc_tables.w160_im[n2 * n1]);     // tw[0] = 1 is not used
    }
    const float2 w16 = make_float2(g_mfcc_tables.w160_re[10 * n2], g_mfcc_tables.w160_im[10 * n2]);   // W_16^n2
    // step 2 (thread = (frame of five, a))
    const int a = lane / 5, fl5 = lane % 5;     // a-major: the slot reads of a half-warp then spread over all banks
    const int binA = mel_bin[lane], binB = mel_bin[na * 32 + lane];    // filters own CONSECUTIVE bins
    __syncthreads();

    const int64_t s0 = pcm_off[u];
    const int64_t n_samples = pcm_off[u + 1] - s0;
    const SampleT* __restrict__ x = pcm + s0;
    // two samples per load when the utterance starts on an even sample (the pair is then naturally aligned)
    const bool pair_ok = ((s0 & 1) == 0) && ((reinterpret_cast<uintptr_t>(pcm) & (2 * sizeof(SampleT) - 1)) == 0);
    using PairT = typename Pair<SampleT>::type;
    float* area = sm.area[warp];
    float vmax = 0.f;

    // raw samples of one step-1 item (frame t, this lane's n2): z[16 n1 + n2] = (x[32 n1 + 2 n2], x[32 n1 + 2 n2 + 1]),
    // zero outside the utterance (centre padding)
    auto fetch = [&](int t, float2* dst) {
        const int64_t base = (int64_t)kHop * t - kHalf;
        if (base >= 0 && base + kNfft <= n_samples) {          // interior frame: no bounds checks
            const SampleT* __restrict__ xb = x + base + 2 * n2;
            if (pair_ok) {
#pragma unroll
                for (int n1 = 0; n1 < 10; ++n1) {
                    const PairT s = __ldg(reinterpret_cast<const PairT*>(xb + 32 * n1));
                    dst[n1] = make_float2(to_f32(s.x), to_f32(s.y));
                }
            } else {
#pragma unroll
                for (int n1 = 0; n1 < 10; ++n1) dst[n1] = make_float2(to_f32(__ldg(xb + 32 * n1)), to_f32(__ldg(xb + 32 * n1 + 1)));
            }
        } else {
#pragma unroll
            for (int n1 = 0; n1 < 10; ++n1) {
                const int64_t i0 = base + 32 * n1 + 2 * n2, i1 = i0 + 1;
                dst[n1] = make_float2((i0 >= 0 && i0 < n_samples) ? to_f32(__ldg(x + i0)) : 0.f,
                                      (i1 >= 0 && i1 < n_samples) ? to_f32(__ldg(x + i1)) : 0.f);
            }
        }
    };
    // the samples of the next step-1 item are requested one pass ahead (across batches too), so that the loads
    // are in flight while the current item is transformed
    float2 vn[10];
#pragma unroll
    for (int n1 = 0; n1 < 10; ++n1) vn[n1] = make_float2(0.f, 0.f);
    if (t_begin + warp * kBatchA + fl < t_end) fetch(t_begin + warp * kBatchA + fl, vn);

    for (int tb = t_begin + warp * kBatchA; tb < t_end; tb += kWarpsA * kBatchA) {
        // pull the samples of this warp's next batch into L2 (11 hops = 1760 + 160 samples, one 128-byte line per lane
        // and round): the register prefetch below then only has to cover an L2 hit
        {
            const int tbn = tb + kWarpsA * kBatchA;
            if (tbn < t_end) {
                const int64_t lo = max((int64_t)0, (int64_t)kHop * tbn - kHalf);
                const int64_t hi = min(n_samples, (int64_t)kHop * (tbn + kBatchA) + kHalf);
                constexpr int kPerLine = 128 / (int)sizeof(SampleT);
                for (int64_t i = lo + (int64_t)lane * kPerLine; i < hi; i += 32 * kPerLine)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(x + i));
            }
        }
        // ---------------- step 1: window, 10-point DFTs, twiddle, store the 11 slots
#pragma unroll 1
        for (int p = 0; p < kBatchA / 2; ++p) {
            const int fb = 2 * p + fl, t = tb + fb;
            float2 v[10];
#pragma unroll
            for (int n1 = 0; n1 < 10; ++n1)           // never contracted: f32 and s16 inputs agree bit for bit
                v[n1] = emul(vn[n1], make_float2(hw[2 * n1], hw[2 * n1 + 1]));
            const int tn = (p < kBatchA / 2 - 1) ? t + 2 : tb + kWarpsA * kBatchA + fl;
            if (tn < t_end) fetch(tn, vn);
            if (t < t_end) {
                float2 y[10];
                dft10(v, y);
                float2* slot = reinterpret_cast<float2*>(area + fb * kFramePitch) + n2;
                slot[0] = y[0];
                slot[10 * kSlotPitch] = cmul(y[0], w16);
#pragma unroll
                for (int k1 = 1; k1 < 10; ++k1) slot[k1 * kSlotPitch] = cmul(y[k1], tw[k1]);
            }
        }
        __syncwarp();
        // ---------------- step 2: two 16-point FFTs per thread, real-input post-pass, power spectrum
#pragma unroll 1
        for (int p = 0; p < (kBatchA + 4) / 5; ++p) {
            const int fb = 5 * p + fl5;
            const bool active = lane < 30 && fb < kBatchA && tb + fb < t_end;
            float* fa = area + fb * kFramePitch;
            float2 za[16], zb[16];
            if (active) {
                const float2* sa = reinterpret_cast<const float2*>(fa) + a * kSlotPitch;
                const float2* sb = reinterpret_cast<const float2*>(fa) + (10 - a) * kSlotPitch;
#pragma unroll
                for (int i = 0; i < 16; ++i) { za[i] = sa[i]; zb[i] = sb[i]; }
            }
            __syncwarp();                                   // the slots of these frames are in registers: their areas take the power spectra
            if (active) {
                fft16(za);
                fft16(zb);
                const float2* wp = sm.wpost + a * kSlotPitch;
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) {
                    // X[k] = (e + w o) / 2, X[160 - k] = conj(e - w o) / 2 with A = Z[k], B = Z[160 - k]
                    const float2 A = za[k2], B = zb[15 - k2];
                    const float2 e = make_float2(A.x + B.x, A.y - B.y);
                    const float2 o = make_float2(A.y + B.y, B.x - A.x);
                    const float2 wo = cmul(wp[k2], o);
                    const float2 xp = cadd(e, wo), xm = csub(e, wo);
                    const int k = a + 10 * k2;
                    fa[k] = 0.25f * (xp.x * xp.x + xp.y * xp.y);
                    fa[kNfft / 2 - k] = 0.25f * (xm.x * xm.x + xm.y * xm.y);
                }
            }
        }
        __syncwarp();
        // ---------------- mel filterbank (lane-balanced table), one frame at a time
        const int nf = min(kBatchA, t_end - tb);
#pragma unroll 1
        for (int fb = 0; fb < nf; ++fb) {
            const float* pw = area + fb * kFramePitch;
            float accA = 0.f, accB = 0.f;
            if (NA > 0) {
#pragma unroll
                for (int it = 0; it < NA; ++it) accA = fmaf(sm.mel_w[it * 32 + lane], pw[binA + it], accA);
#pragma unroll
                for (int it = 0; it < NB; ++it) accB = fmaf(sm.mel_w[(NA + it) * 32 + lane], pw[binB + 4 * it], accB);
            } else {
                for (int it = 0; it < na; ++it) accA = fmaf(sm.mel_w[it * 32 + lane], pw[binA + it], accA);
                for (int it = 0; it < nb; ++it) accB = fmaf(sm.mel_w[(na + it) * 32 + lane], pw[binB + 4 * it], accB);
            }
            accB += __shfl_xor_sync(FULL, accB, 1);
            accB += __shfl_xor_sync(FULL, accB, 2);
            float* mo = mel_out + (f0 + tb + fb) * kMels;
            mo[lane] = accA;
            if ((lane & 3) == 0) mo[32 + (lane >> 2)] = accB;
            vmax = fmaxf(vmax, fmaxf(accA, accB));
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(FULL, vmax, o));
    if (lane == 0) atomicMax(reinterpret_cast<int*>(utt_max + u), __float_as_int(vmax));   // mel >= 0
}

// ------------------------------------------------------------------------------------------
// Kernel A, real-input-first form (the 16 kHz table of every reference call site: na = 11, nb = 5)
// ------------------------------------------------------------------------------------------
// The 320-point real DFT is split 20 x 16 with the REAL transform first, so that no real-input post-pass exists:
//   n = 16 n1 + n2, k = k1 + 20 k2
//   step 1, thread (frame, n2):  T[k1][n2] = W_320^(n2 k1) * sum_n1 w[n] x[n] W_20^(n1 k1)   for k1 = 0..10 only (real input:
//            k1 = 11..19 are conjugates).  The 20-point real DFT is a prime-factor 4 x 5 transform without twiddles:
//            n1 = (5 a + 4 b) mod 20, k1 = c mod 4 = d mod 5; five real 4-point DFTs with the window folded into their
//            first butterflies, then real 5-point DFTs for c = 0 and c = 2 and one complex 5-point DFT for c = 1
//            (c = 3 is its conjugate).
//   step 2, thread (frame, k1):  X[k1 + 20 k2] = sum_n2 T[k1][n2] W_16^(n2 k2): one 16-point FFT, whose 16 outputs are 16
//            DIFFERENT bins of the power spectrum (bins above 160 are the mirror images 320 - k); the rows k1 = 0 and
//            k1 = 10 mirror onto themselves and write 9 and 8 distinct bins.
// A warp works on batches of 8 frames: step 1 in 4 passes of 2 frames x 16 lanes, step 2 in 3 passes of 8 frames x 4 rows,
// then the filterbank for all 8 frames at once with a filter per lane: the power spectra are stored [plane of 4 frames][bin]
// [frame], so one 16-byte load brings a bin of four frames and a weight is fetched once per 8 frames.
// Shared memory per warp: 11 rows x 8 frames x (128 + 16) bytes; the two power planes take the place of rows 4..10, which is
// why step 2 starts with the rows 8..10 (results kept in registers), then 4..7, then 0..3.
namespace r20 {
constexpr int kWarps = 4;
constexpr int kBatch = 8;
constexpr int kFrameB = 144;                 // bytes per (row, frame): 16 complex + 16: the 16-byte loads of 8 frames hit 8 bank groups
constexpr int kRowB = kBatch * kFrameB;      // 1152
constexpr int kRows = 11;
constexpr int kAreaB = kRows * kRowB;        // 12672 bytes per warp
constexpr int kPlaneB = 3136;                // 196 bins x 16 bytes; 16 banks mod 32: the two planes' stores never collide
constexpr int kPlane0 = 4 * kRowB;
constexpr int kNA = 11, kNB = 5;
static_assert(kPlane0 + 2 * kPlaneB <= kAreaB, "power planes must fit the rows 4..10");
static_assert(kBins * 16 <= kPlaneB, "plane too small");

struct __align__(16) Smem {
    unsigned char area[kWarps][kAreaB];
    float mel_w[(kNA + kNB) * 32];
};

// real-input 5-point DFT: V[0] = v0, V[1] = m1 - i q1, V[4] = m1 + i q1, V[2] = m2 - i q2, V[3] = m2 + i q2
__device__ __forceinline__ void rdft5(float r0, float r1, float r2, float r3, float r4,
                                      float& v0, float& m1, float& q1, float& m2, float& q2) {
    const float C1 = 0.30901699437494745f, C2 = -0.80901699437494745f;
    const float S1 = 0.95105651629515353f, S2 = 0.58778525229247314f;
    const float t1 = r1 + r4, t2 = r2 + r3, t3 = r1 - r4, t4 = r2 - r3;
    v0 = (r0 + t1) + t2;
    m1 = fmaf(C2, t2, fmaf(C1, t1, r0));
    m2 = fmaf(C1, t2, fmaf(C2, t1, r0));
    q1 = fmaf(S2, t4, S1 * t3);
    q2 = fmaf(-S1, t4, S2 * t3);
}
}  // namespace r20

template <typename SampleT>
__global__ void __launch_bounds__(r20::kWarps * 32, LOE_MELR_MINB)
mfcc_mel_r_kernel(const SampleT* __restrict__ pcm, const int64_t* __restrict__ pcm_off,
                  const int64_t* __restrict__ frm_off, const int32_t* __restrict__ mel_bin,
                  const float* __restrict__ mel_w, int chunk,
                  float* __restrict__ mel_out, float* __restrict__ utt_max) {
    using namespace r20;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int u = blockIdx.x;
    const int64_t f0 = frm_off[u];
    const int T = (int)(frm_off[u + 1] - f0);
    const int t_begin = blockIdx.y * chunk;
    if (t_begin >= T) return;
    const int t_end = min(T, t_begin + chunk);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr unsigned FULL = 0xffffffffu;

    for (int i = tid; i < (kNA + kNB) * 32; i += kWarps * 32) sm.mel_w[i] = mel_w[i];
    // power planes start out finite (zero-weight table entries read bins of frames that do not exist in a partial batch)
    for (int i = tid; i < kWarps * kAreaB / 4; i += kWarps * 32) reinterpret_cast<float*>(&sm.area[0][0])[i] = 0.f;

    // step 1 constants, thread = (frame parity fl, n2): window taps w[16 n1 + n2] and W_320^(n2 k1), k1 = 1..10
    const int n2 = lane & 15, fl = lane >> 4;
    float hw[20];
    float2 tw[11];
#pragma unroll
    for (int n1 = 0; n1 < 20; ++n1) hw[n1] = g_mfcc_tables.hann[16 * n1 + n2];
#pragma unroll
    for (int k1 = 1; k1 <= 10; ++k1) tw[k1] = make_float2(g_mfcc_tables.w320_re[n2 * k1], g_mfcc_tables.w320_im[n2 * k1]);
    // step 2: thread = (frame f8, row kq + 4 * pass)
    const int f8 = lane & 7, kq = lane >> 3;
    const int binA = mel_bin[lane], binB = mel_bin[kNA * 32 + lane];
    __syncthreads();

    const int64_t s0 = pcm_off[u];
    const int64_t n_samples = pcm_off[u + 1] - s0;
    const SampleT* __restrict__ x = pcm + s0;
    unsigned char* area = sm.area[warp];
    float vmax = 0.f;

    const int ns = (int)n_samples;                    // one utterance: fits 32 bits
    // samples of one step-1 item: x[160 t - 160 + 16 n1 + n2], zero outside the utterance (centre padding).
    // Interior frames (all but the first and the last one or two of an utterance) are requested WITHOUT any test, from a
    // base clamped into the utterance, as soon as the window butterflies have consumed the current samples (into the
    // same registers: the loads are in flight for the rest of the item, across batches too).  A frame that reaches
    // outside the utterance is fetched again, with bounds tests, when its turn comes.
    const int base_max = ns - kNfft;                  // >= 0: an utterance has at least 9 frames
    auto edge = [&](int t) { const int base = kHop * t - kHalf; return base < 0 || base > base_max; };
    auto fetch_interior = [&](int t, float* dst) {
        const SampleT* __restrict__ xb = x + min(max(kHop * t - kHalf, 0), base_max) + n2;
#pragma unroll
        for (int n1 = 0; n1 < 20; ++n1) dst[n1] = to_f32(__ldg(xb + 16 * n1));
    };
    auto fetch_edge = [&](int t, float* dst) {
        const int base = kHop * t - kHalf;
#pragma unroll
        for (int n1 = 0; n1 < 20; ++n1) {
            const int i = base + 16 * n1 + n2;
            dst[n1] = (i >= 0 && i < ns) ? to_f32(__ldg(x + i)) : 0.f;
        }
    };
    float v[20];
    fetch_interior(t_begin + warp * kBatch + fl, v);

    for (int tb = t_begin + warp * kBatch; tb < t_end; tb += kWarps * kBatch) {
        {   // pull the samples of this warp's next batch into L2
            const int tbn = tb + kWarps * kBatch;
            if (tbn < t_end) {
                constexpr int kPerLine = 128 / (int)sizeof(SampleT);
                const int lo = max(0, kHop * tbn - kHalf), hi = min(ns, kHop * (tbn + kBatch) + kHalf);
                for (int i = lo + lane * kPerLine; i < hi; i += 32 * kPerLine)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(x + i));
            }
        }
        // ---------------- step 1
#pragma unroll 1
        for (int p = 0; p < kBatch / 2; ++p) {
            const int fb = 2 * p + fl, t = tb + fb;
            if (edge(t) && t < t_end) fetch_edge(t, v);
            // five real 4-point DFTs over a (n1 = (5 a + 4 b) mod 20), window folded into the first butterflies
            float u0[5], u2[5];
            float2 u1[5];
#pragma unroll
            for (int b = 0; b < 5; ++b) {
                const int i0 = (4 * b) % 20, i1 = (5 + 4 * b) % 20, i2 = (10 + 4 * b) % 20, i3 = (15 + 4 * b) % 20;
                const float p0 = hw[i0] * v[i0], p1 = hw[i1] * v[i1];
                const float s0_ = fmaf(hw[i2], v[i2], p0), s1_ = fmaf(-hw[i2], v[i2], p0);
                const float s2_ = fmaf(hw[i3], v[i3], p1), s3_ = fmaf(-hw[i3], v[i3], p1);
                u0[b] = s0_ + s2_;
                u2[b] = s0_ - s2_;
                u1[b] = make_float2(s1_, -s3_);
            }
            fetch_interior((p < kBatch / 2 - 1) ? t + 2 : tb + kWarps * kBatch + fl, v);
            if (t < t_end) {
                float2 Y[11];
                {
                    float v0, m1, q1, m2, q2;
                    rdft5(u0[0], u0[1], u0[2], u0[3], u0[4], v0, m1, q1, m2, q2);      // c = 0: k1 = 0, 4 (d = 4), 8 (d = 3)
                    Y[0] = make_float2(v0, 0.f); Y[4] = make_float2(m1, q1); Y[8] = make_float2(m2, q2);
                    rdft5(u2[0], u2[1], u2[2], u2[3], u2[4], v0, m1, q1, m2, q2);      // c = 2: k1 = 10, 6 (d = 1), 2 (d = 2)
                    Y[10] = make_float2(v0, 0.f); Y[6] = make_float2(m1, -q1); Y[2] = make_float2(m2, -q2);
                    float2 V[5];
                    dft5(u1[0], u1[1], u1[2], u1[3], u1[4], V);                           // c = 1: k1 = 5, 1, 17, 13, 9
                    Y[5] = V[0]; Y[1] = V[1]; Y[9] = V[4];
                    Y[3] = make_float2(V[2].x, -V[2].y);                                   // conj of k1 = 17
                    Y[7] = make_float2(V[3].x, -V[3].y);                                   // conj of k1 = 13
                }
                float2* slot = reinterpret_cast<float2*>(area + fb * kFrameB) + n2;
                slot[0] = Y[0];
#pragma unroll
                for (int k1 = 1; k1 < 10; ++k1) slot[k1 * (kRowB / 8)] = cmul(Y[k1], tw[k1]);
                slot[10 * (kRowB / 8)] = make_float2(Y[10].x * tw[10].x, Y[10].x * tw[10].y);
            }
        }
        __syncwarp();
        // ---------------- step 2
        const bool fvalid = tb + f8 < t_end;
        float* plane = reinterpret_cast<float*>(area + kPlane0 + (f8 >> 2) * kPlaneB) + (f8 & 3);
        auto load_row = [&](int k1, float2* z) {
            const float4* src = reinterpret_cast<const float4*>(area + k1 * kRowB + f8 * kFrameB);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 q = src[c];
                z[2 * c] = make_float2(q.x, q.y);
                z[2 * c + 1] = make_float2(q.z, q.w);
            }
        };
        auto store_pw = [&](int k1, const float* pw) {
            float* lo = plane + 4 * k1;                    // bin k1 + 20 k2, k2 = 0..7
            float* hi = plane + 4 * (320 - k1);            // bin 320 - k1 - 20 k2, k2 = 8..15
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) lo[80 * k2] = pw[k2];
#pragma unroll
            for (int k2 = 8; k2 < 16; ++k2) hi[-80 * k2] = pw[k2];
        };
        float pc[16];
        {
            float2 z[16];
            const bool act = fvalid && kq < 3;
            if (act) {
                load_row(8 + kq, z);
                fft16(z);
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) pc[k2] = fmaf(z[k2].x, z[k2].x, z[k2].y * z[k2].y);
            }
            float2 zb[16];
            if (fvalid) load_row(4 + kq, zb);
            __syncwarp();                                   // rows 4..10 are in registers: their area takes the power planes
            if (act) store_pw(8 + kq, pc);
            if (fvalid) {
                fft16(zb);
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) pc[k2] = fmaf(zb[k2].x, zb[k2].x, zb[k2].y * zb[k2].y);
                store_pw(4 + kq, pc);
                load_row(kq, z);
                fft16(z);
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) pc[k2] = fmaf(z[k2].x, z[k2].x, z[k2].y * z[k2].y);
                store_pw(kq, pc);
            }
        }
        __syncwarp();
        // ---------------- mel filterbank, 8 frames at once
        {
            const int nf = min(kBatch, t_end - tb);
            const unsigned char* pa = area + kPlane0 + binA * 16;
            float acc[8];
#pragma unroll
            for (int f = 0; f < 8; ++f) acc[f] = 0.f;
#pragma unroll
            for (int it = 0; it < kNA; ++it) {
                const float w = sm.mel_w[it * 32 + lane];
                const float4 p = *reinterpret_cast<const float4*>(pa + it * 16);
                const float4 q = *reinterpret_cast<const float4*>(pa + kPlaneB + it * 16);
                acc[0] = fmaf(w, p.x, acc[0]); acc[1] = fmaf(w, p.y, acc[1]); acc[2] = fmaf(w, p.z, acc[2]); acc[3] = fmaf(w, p.w, acc[3]);
                acc[4] = fmaf(w, q.x, acc[4]); acc[5] = fmaf(w, q.y, acc[5]); acc[6] = fmaf(w, q.z, acc[6]); acc[7] = fmaf(w, q.w, acc[7]);
            }
            float* mo = mel_out + (f0 + tb) * kMels + lane;
#pragma unroll
            for (int f = 0; f < 8; ++f)
                if (f < nf) { mo[f * kMels] = acc[f]; vmax = fmaxf(vmax, acc[f]); }
            const unsigned char* pb = area + kPlane0 + binB * 16;
#pragma unroll
            for (int f = 0; f < 8; ++f) acc[f] = 0.f;
#pragma unroll
            for (int it = 0; it < kNB; ++it) {
                const float w = sm.mel_w[(kNA + it) * 32 + lane];
                const float4 p = *reinterpret_cast<const float4*>(pb + it * 64);
                const float4 q = *reinterpret_cast<const float4*>(pb + kPlaneB + it * 64);
                acc[0] = fmaf(w, p.x, acc[0]); acc[1] = fmaf(w, p.y, acc[1]); acc[2] = fmaf(w, p.z, acc[2]); acc[3] = fmaf(w, p.w, acc[3]);
                acc[4] = fmaf(w, q.x, acc[4]); acc[5] = fmaf(w, q.y, acc[5]); acc[6] = fmaf(w, q.z, acc[6]); acc[7] = fmaf(w, q.w, acc[7]);
            }
#pragma unroll
            for (int f = 0; f < 8; ++f) {
                acc[f] += __shfl_xor_sync(FULL, acc[f], 1);
                acc[f] += __shfl_xor_sync(FULL, acc[f], 2);
            }
            float* mb = mel_out + (f0 + tb) * kMels + 32 + (lane >> 2);
            if ((lane & 3) == 0) {
#pragma unroll
                for (int f = 0; f < 8; ++f)
                    if (f < nf) { mb[f * kMels] = acc[f]; vmax = fmaxf(vmax, acc[f]); }
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(FULL, vmax, o));
    if (lane == 0) atomicMax(reinterpret_cast<int*>(utt_max + u), __float_as_int(vmax));   // mel >= 0
}

// ------------------------------------------------------------------------------------------
// Kernel B
// ------------------------------------------------------------------------------------------
// One thread per frame.  A CTA takes 120 output frames of one utterance plus the 8 halo frames the width-9
// delta filters reach: 128 rows of mel energies are converted to dB while they are copied (coalesced) into
// shared memory; each thread then pulls its row into registers and runs the 13 x 40 DCT against the constant
// bank (the coefficient is an immediate-like operand of the FMA: no load instruction), normalises the static
// block in registers, and the deltas read the cepstra of the neighbouring rows from shared memory.
constexpr int kTileB = 120;
constexpr int kRowsB = kTileB + 8;      // = threads per CTA
constexpr int kMelPitch = kMels + 1;    // odd pitches: row-per-lane accesses are bank-conflict free
constexpr int kCepPitch = kCeps;
constexpr int kOutPitch = kFeat;


__global__ void __launch_bounds__(kRowsB)
mfcc_ceps_kernel(const float* __restrict__ mel, const float* __restrict__ utt_max,
                 const int64_t* __restrict__ frm_off, float* __restrict__ feat, uint8_t* __restrict__ a_img, float* __restrict__ inv2) {
    __shared__ float s_lm[kRowsB * kMelPitch];  // dB mel rows; reused for the output tile (kTileB * 39 floats)
    __shared__ float s_c[kRowsB * kCepPitch];
    const int u = blockIdx.x;
    const int64_t f0 = frm_off[u];
    const int T = (int)(frm_off[u + 1] - f0);
    const int t0 = blockIdx.y * kTileB;
    if (t0 >= T) return;
    const int t1 = min(T, t0 + kTileB);
    const int tid = threadIdx.x;
    const int g0 = max(0, min(t0 - 4, T - 9));
    const int g1 = min(T - 1, max(t1 + 3, 8));
    const int ng = g1 - g0 + 1;                 // <= kRowsB

    // 10 log10(x) = kDbPerLog2 * log2(x); MUFU.LG2 (__log2f, 2 ulp) instead of the 20-instruction log10f
    constexpr float kDbPerLog2 = 3.0102999566398120f;
    const float ref_db = kDbPerLog2 * __log2f(fmaxf(1e-10f, utt_max[u]));
    // mel rows are 160 bytes: four energies per 16-byte load (one row / column split per load instead of per value)
    const float4* __restrict__ src4 = reinterpret_cast<const float4*>(mel + (f0 + g0) * kMels);
    static_assert(kMels % 4 == 0, "vector loads of the mel rows");
    // all of a thread's loads are requested before the first logarithm (128 rows x 10 pieces = 10 per thread)
    constexpr int kPiecesB = kMels / 4;
    float4 e[kPiecesB];
#pragma unroll
    for (int r = 0; r < kPiecesB; ++r) {
        const int i = tid + r * kRowsB;
        e[r] = (i < ng * kPiecesB) ? __ldg(src4 + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    }
#pragma unroll
    for (int r = 0; r < kPiecesB; ++r) {
        const int i = tid + r * kRowsB;
        if (i < ng * kPiecesB) {
            const int j = i / kPiecesB, m = 4 * (i - j * kPiecesB);
            float* o = s_lm + j * kMelPitch + m;
            o[0] = fmaxf(fmaf(kDbPerLog2, __log2f(fmaxf(1e-10f, e[r].x)), -ref_db), -80.0f);
            o[1] = fmaxf(fmaf(kDbPerLog2, __log2f(fmaxf(1e-10f, e[r].y)), -ref_db), -80.0f);
            o[2] = fmaxf(fmaf(kDbPerLog2, __log2f(fmaxf(1e-10f, e[r].z)), -ref_db), -80.0f);
            o[3] = fmaxf(fmaf(kDbPerLog2, __log2f(fmaxf(1e-10f, e[r].w)), -ref_db), -80.0f);
        }
    }
    __syncthreads();
    float c[kCeps];
    if (tid < ng) {
        float lm[kMels];
#pragma unroll
        for (int m = 0; m < kMels; ++m) lm[m] = s_lm[tid * kMelPitch + m];
        // DCT-II symmetry: cos((2 (39 - m) + 1) k pi / 80) = (-1)^k cos((2 m + 1) k pi / 80), so even rows see
        // lm[m] + lm[39 - m] and odd rows lm[m] - lm[39 - m]: 20 FMAs per coefficient instead of 40
        float ev[kMels / 2], od[kMels / 2];
#pragma unroll
        for (int m = 0; m < kMels / 2; ++m) { ev[m] = lm[m] + lm[kMels - 1 - m]; od[m] = lm[m] - lm[kMels - 1 - m]; }
#pragma unroll
        for (int k = 0; k < kCeps; ++k) {
            float acc = 0.f;
#pragma unroll
            for (int m = 0; m < kMels / 2; ++m) acc = fmaf(c_dct[k * kMels + m], (k & 1) ? od[m] : ev[m], acc);
            c[k] = acc;
            s_c[tid * kCepPitch + k] = acc;
        }
    }
    __syncthreads();                            // cepstra visible; every read of s_lm is done
    float* s_out = s_lm;
    const int r = tid - (t0 - g0);              // output frame handled by this thread: t0 + r
    if (r >= 0 && r < t1 - t0) {
        float mean = 0.f;
#pragma unroll
        for (int k = 0; k < kCeps; ++k) mean += c[k];
        mean *= (1.0f / kCeps);
        float var = 0.f;
#pragma unroll
        for (int k = 0; k < kCeps; ++k) { const float d = c[k] - mean; var = fmaf(d, d, var); }
        const float inv = 1.0f / (sqrtf(var * (1.0f / kCeps)) + 1e-8f);
        float* o = s_out + r * kOutPitch;
#pragma unroll
        for (int k = 0; k < kCeps; ++k) o[k] = (c[k] - mean) * inv;
        const int cc = min(max(t0 + r, 4), T - 5) - g0;      // centre row of the delta window
        float d1[kCeps], d2[kCeps];
#pragma unroll
        for (int k = 0; k < kCeps; ++k) { d1[k] = 0.f; d2[k] = 0.f; }
#pragma unroll
        for (int q = -4; q <= 4; ++q) {
            const float* row = s_c + (cc + q) * kCepPitch;
#pragma unroll
            for (int k = 0; k < kCeps; ++k) {
                const float cv = row[k];
                d1[k] = fmaf(c_sg1[q + 4], cv, d1[k]);
                d2[k] = fmaf(c_sg2[q + 4], cv, d2[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < kCeps; ++k) { o[kCeps + k] = d1[k]; o[2 * kCeps + k] = d2[k]; }
    }
    __syncthreads();
    const int nt = t1 - t0;
    if (feat) {
        float* __restrict__ dst = feat + (f0 + t0) * kFeat;
        for (int i = tid; i < nt * kFeat; i += kRowsB) dst[i] = s_out[i];
    }
    // optional second output: the row as the binary16 hi / lo A operand of the 3xFP16 emission kernel (h16_stage.cuh), in
    // the tile-major image that kernel bulk-copies -- tile = global frame / 128, 16 bytes per row and chunk: consecutive
    // threads write consecutive 16-byte pieces
    if (a_img && tid < nt) {
        const int64_t f = f0 + t0 + tid;
        uint8_t* a_row = a_img + (size_t)(f >> 7) * h16::kImgTileBytes + (size_t)(f & 127) * 16;
        inv2[f] = h16::stage_row(s_out + tid * kOutPitch, a_row);
    }
}

}  // namespace loe

static int mfcc_launch(const void* pcm_dev, int pcm_format, const int64_t* pcm_off_dev, const int64_t* frm_off_dev,
                       int n_utt, int64_t total_frames, int max_frames, int min_frames,
                       const int32_t* mel_bin_dev, const float* mel_w_dev, int mel_na, int mel_nb,
                       float* mel_ws_dev, float* utt_max_dev, float* feat_dev, void* stream, int phases,
                       void* a_img_dev = nullptr, float* inv2_dev = nullptr) {
    using namespace loe;
    if (n_utt <= 0 || total_frames <= 0) return LOE_OK;
    if (min_frames < 9) {
        set_error("MFCC needs at least 9 frames per utterance for the width-9 delta filter (got %d)", min_frames);
        return LOE_ERR_VALUE;
    }
    if (mel_na < 0 || mel_nb < 0 || mel_na > LOE_MEL_NA_MAX || mel_nb > LOE_MEL_NB_MAX) {
        set_error("mel table iteration counts out of range (na=%d, nb=%d)", mel_na, mel_nb);
        return LOE_ERR_VALUE;
    }
    if ((reinterpret_cast<uintptr_t>(mel_ws_dev) & 15) != 0) {
        set_error("mel_ws_dev must be 16-byte aligned (the cepstrum kernel reads it with 16-byte loads)");
        return LOE_ERR_VALUE;
    }
    int st = ensure_tables();
    if (st != LOE_OK) return st;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (phases & 1) {
        LOE_CUDA(cudaMemsetAsync(utt_max_dev, 0, sizeof(float) * (size_t)n_utt, s));
        // frames per CTA: as many as possible (the per-CTA set-up -- constants, zeroed shared memory, first loads -- is
        // paid once) while the batch still fills the machine a few times over: about 12 CTAs per SM in flight or queued
        int sms = 0, dev = 0;
        LOE_CUDA(cudaGetDevice(&dev));
        LOE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const bool k16 = (mel_na == r20::kNA && mel_nb == r20::kNB);      // the 16 kHz table of every reference call site
        const int round = k16 ? r20::kWarps * r20::kBatch : kWarpsA * kBatchA;    // whole rounds of the CTA's warps
        int64_t chunk64 = total_frames / (12 * (int64_t)sms);
        chunk64 = ((chunk64 + round - 1) / round) * round;
        const int chunk = (int)(chunk64 < kMinChunkA ? kMinChunkA : chunk64 > (1 << 20) ? (1 << 20) : chunk64);
        dim3 ga((unsigned)n_utt, (unsigned)((max_frames + chunk - 1) / chunk));
        // the kernels' shared memory (exchange area of 4 warps) exceeds the 48 KB default: opt in
#define LOE_MEL_LAUNCH(T)                                                                                                \
        LOE_CUDA(cudaFuncSetAttribute(mfcc_mel_kernel<T, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemA))); \
        mfcc_mel_kernel<T, 0, 0><<<ga, kWarpsA * 32, sizeof(SmemA), s>>>((const T*)pcm_dev, pcm_off_dev, frm_off_dev, mel_bin_dev, \
                                                                        mel_w_dev, mel_na, mel_nb, chunk, mel_ws_dev, utt_max_dev)
#define LOE_MELR_LAUNCH(T)                                                                                               \
        LOE_CUDA(cudaFuncSetAttribute(mfcc_mel_r_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(r20::Smem))); \
        mfcc_mel_r_kernel<T><<<ga, r20::kWarps * 32, sizeof(r20::Smem), s>>>((const T*)pcm_dev, pcm_off_dev, frm_off_dev, mel_bin_dev, \
                                                                            mel_w_dev, chunk, mel_ws_dev, utt_max_dev)
        if (pcm_format == LOE_PCM_F32) { if (k16) { LOE_MELR_LAUNCH(float); } else { LOE_MEL_LAUNCH(float); } }
        else if (pcm_format == LOE_PCM_S16) { if (k16) { LOE_MELR_LAUNCH(short); } else { LOE_MEL_LAUNCH(short); } }
        else { set_error("unknown pcm_format %d", pcm_format); return LOE_ERR_VALUE; }
#undef LOE_MEL_LAUNCH
#undef LOE_MELR_LAUNCH
        LOE_LAUNCH_CHECK("mfcc_mel_kernel");
    }
    if (phases & 2) {
        dim3 gb((unsigned)n_utt, (unsigned)((max_frames + kTileB - 1) / kTileB));
        if (a_img_dev) {
            // rows of the last image tile beyond the batch: zero operand, unit scale (the kernel writes the rows that exist)
            const int64_t full = (total_frames / 128) * 128;
            if (full < total_frames) {
                LOE_CUDA(cudaMemsetAsync(static_cast<uint8_t*>(a_img_dev) + (size_t)(total_frames / 128) * h16::kImgTileBytes, 0, h16::kImgTileBytes, s));
                LOE_CUDA(cudaMemsetAsync(inv2_dev + full, 0, sizeof(float) * 128, s));
            }
        }
        mfcc_ceps_kernel<<<gb, kRowsB, 0, s>>>(mel_ws_dev, utt_max_dev, frm_off_dev, feat_dev, static_cast<uint8_t*>(a_img_dev), inv2_dev);
        LOE_LAUNCH_CHECK("mfcc_ceps_kernel");
    }
    return LOE_OK;
}

extern "C" int loe_mfcc_dev(const void* pcm_dev, int pcm_format, const int64_t* pcm_off_dev, const int64_t* frm_off_dev,
                            int n_utt, int64_t total_frames, int max_frames, int min_frames,
                            const int32_t* mel_bin_dev, const float* mel_w_dev, int mel_na, int mel_nb,
                            float* mel_ws_dev, float* utt_max_dev, float* feat_dev, void* stream) {
    return mfcc_launch(pcm_dev, pcm_format, pcm_off_dev, frm_off_dev, n_utt, total_frames, max_frames, min_frames, mel_bin_dev,
                       mel_w_dev, mel_na, mel_nb, mel_ws_dev, utt_max_dev, feat_dev, stream, 3);
}

extern "C" int loe_mfcc_img_dev(const void* pcm_dev, int pcm_format, const int64_t* pcm_off_dev, const int64_t* frm_off_dev,
                                int n_utt, int64_t total_frames, int max_frames, int min_frames,
                                const int32_t* mel_bin_dev, const float* mel_w_dev, int mel_na, int mel_nb,
                                float* mel_ws_dev, float* utt_max_dev, float* feat_dev, void* a_img_dev, float* inv2_dev, void* stream, int phases) {
    using namespace loe;
    if (!a_img_dev || !inv2_dev) { set_error("a_img_dev / inv2_dev required"); return LOE_ERR_VALUE; }
    if ((reinterpret_cast<uintptr_t>(a_img_dev) & 15) != 0) { set_error("a_img_dev must be 16-byte aligned"); return LOE_ERR_VALUE; }
    return mfcc_launch(pcm_dev, pcm_format, pcm_off_dev, frm_off_dev, n_utt, total_frames, max_frames, min_frames, mel_bin_dev,
                       mel_w_dev, mel_na, mel_nb, mel_ws_dev, utt_max_dev, feat_dev, stream, phases, a_img_dev, inv2_dev);
}

extern "C" int loe_mfcc_phase_dev(const void* pcm_dev, int pcm_format, const int64_t* pcm_off_dev, const int64_t* frm_off_dev,
                                  int n_utt, int64_t total_frames, int max_frames, int min_frames,
                                  const int32_t* mel_bin_dev, const float* mel_w_dev, int mel_na, int mel_nb,
                                  float* mel_ws_dev, float* utt_max_dev, float* feat_dev, void* stream, int phases) {
    return mfcc_launch(pcm_dev, pcm_format, pcm_off_dev, frm_off_dev, n_utt, total_frames, max_frames, min_frames, mel_bin_dev,
                       mel_w_dev, mel_na, mel_nb, mel_ws_dev, utt_max_dev, feat_dev, stream, phases);
}
