// MFCC front end for sm_100a.  Replaces mfcc.py:24-84 of the reference (librosa pipeline).
//
// Kernel A (mel):   PCM -> Hann window -> 320-point real FFT (one warp per frame: radix-5 in
//                   registers x 32-point shuffle FFT across lanes, real-input post-pass in
//                   shared memory) -> |.|^2 -> sparse slaney filterbank -> mel energies
//                   [frames, 40] + per-utterance maximum (atomicMax on the float bits).
// Kernel B (ceps):  mel -> dB relative to the utterance maximum, floor at -80 dB -> DCT-II
//                   (ortho) 13 ceps -> Savitzky-Golay delta / delta-delta (width 9, edge
//                   frames take the value of the nearest full window) -> per-frame
//                   normalisation of the static block -> features [frames, 39].
// The split is forced by power_to_db(ref=np.max): every frame needs the maximum over the whole
// utterance (mfcc.py:35).  Algorithmic HBM bytes per frame: 640 (PCM) + 156 (features).
#include "common.cuh"
#include <math.h>

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
    float tw160_re[5 * 32], tw160_im[5 * 32];   // W_160^(lane*k1)
    float w32_re[16], w32_im[16];       // W_32^m
    float w320_re[kBins], w320_im[kBins];   // W_320^k
    float dct[kCeps * kMels];           // ortho DCT-II rows
};

__device__ MfccTables g_mfcc_tables;
static bool g_tables_ready[64] = {false};

static int ensure_tables() {
    int dev = 0;
    LOE_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && g_tables_ready[dev]) return LOE_OK;
    static MfccTables h;
    const double PI = 3.14159265358979323846;
    for (int n = 0; n < kNfft; ++n) h.hann[n] = (float)(0.5 - 0.5 * cos(2.0 * PI * n / kNfft));
    for (int k1 = 0; k1 < 5; ++k1)
        for (int l = 0; l < 32; ++l) {
            double a = -2.0 * PI * (double)(l * k1) / kHalf;
            h.tw160_re[k1 * 32 + l] = (float)cos(a);
            h.tw160_im[k1 * 32 + l] = (float)sin(a);
        }
    for (int m = 0; m < 16; ++m) {
        double a = -2.0 * PI * m / 32.0;
        h.w32_re[m] = (float)cos(a);
        h.w32_im[m] = (float)sin(a);
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
    if (dev < 64) g_tables_ready[dev] = true;
    return LOE_OK;
}

// ------------------------------------------------------------------------------------------
// Kernel A
// ------------------------------------------------------------------------------------------
constexpr int kWarpsA = 8;
constexpr int kFramesPerBlockA = 64;
constexpr int kMelItMax = LOE_MEL_NA_MAX + LOE_MEL_NB_MAX;

struct __align__(16) SmemA {
    float mel_w[kMelItMax * 32];
    int mel_bin[kMelItMax * 32];
    float zre[kWarpsA][kHalf];          // per-warp complex spectrum of the packed sequence (split planes:
    float zim[kWarpsA][kHalf];          // the stride-5 transposing store is then bank-conflict free)
    float pw[kWarpsA][kBins + 3 + LOE_MEL_NA_MAX + 4 * LOE_MEL_NB_MAX];   // per-warp power spectrum (+ slack: zero-weight
                                        // table entries may point past the last bin)
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Mel filterbank as a lane-balanced table (host-built, see mfcc.py:mel_lane_tables):
//   round A, iterations [0, na):      lane l accumulates filter l          (filters 0..31)
//   round B, iterations [na, na+nb):  lanes 4q..4q+3 share filter 32 + q   (the 8 widest filters),
//                                     combined with two xor shuffles
// entry (it, lane) = weight mel_w[it*32+lane] applied to power bin mel_bin[it*32+lane] (0-weight padding).
template <typename SampleT, int NA, int NB>
__global__ void __launch_bounds__(kWarpsA * 32)
mfcc_mel_kernel(const SampleT* __restrict__ pcm, const int64_t* __restrict__ pcm_off,
                const int64_t* __restrict__ frm_off, const int32_t* __restrict__ mel_bin,
                const float* __restrict__ mel_w, int na_rt, int nb_rt,
                float* __restrict__ mel_out, float* __restrict__ utt_max) {
    const int na = NA > 0 ? NA : na_rt;                  // NA, NB > 0: compile-time trip counts (loops unroll)
    const int nb = NA > 0 ? NB : nb_rt;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemA& sm = *reinterpret_cast<SmemA*>(smem_raw);
    const int u = blockIdx.x;
    const int64_t f0 = frm_off[u];
    const int T = (int)(frm_off[u + 1] - f0);
    const int t_begin = blockIdx.y * kFramesPerBlockA;
    if (t_begin >= T) return;
    const int t_end = min(T, t_begin + kFramesPerBlockA);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr unsigned FULL = 0xffffffffu;

    for (int i = tid; i < (na + nb) * 32; i += blockDim.x) { sm.mel_w[i] = mel_w[i]; sm.mel_bin[i] = mel_bin[i]; }
    // per-lane constants, fixed for every frame: window taps, 5x32 twiddles, shuffle-FFT twiddles,
    // real-input post-pass twiddles
    float hw[10];
    float2 tw[5], ws[4], wp[5];
#pragma unroll
    for (int n1 = 0; n1 < 5; ++n1) {
        const int n = 2 * (32 * n1 + lane);
        hw[2 * n1] = g_mfcc_tables.hann[n];
        hw[2 * n1 + 1] = g_mfcc_tables.hann[n + 1];
        tw[n1] = make_float2(g_mfcc_tables.tw160_re[n1 * 32 + lane], g_mfcc_tables.tw160_im[n1 * 32 + lane]);
        const int k = lane + 32 * n1;
        wp[n1] = make_float2(g_mfcc_tables.w320_re[k], g_mfcc_tables.w320_im[k]);
    }
    // stage h = 16, 8, 4, 2: lanes with bit h set multiply by W_{2h}^(lane & (h-1)), the others by 1;
    // sg = -1 for the upper lane of a butterfly (o - y), +1 for the lower (y + o)
    float sg[5];
#pragma unroll
    for (int st = 0; st < 5; ++st) {
        const int h = 16 >> st;
        const bool upper = (lane & h) != 0;
        sg[st] = upper ? -1.f : 1.f;
        if (st < 4) {
            const int j = (lane & (h - 1)) * (16 / h);
            ws[st] = upper ? make_float2(g_mfcc_tables.w32_re[j], g_mfcc_tables.w32_im[j]) : make_float2(1.f, 0.f);
        }
    }
    __syncthreads();

    const int64_t s0 = pcm_off[u];
    const int64_t n_samples = pcm_off[u + 1] - s0;
    const SampleT* __restrict__ x = pcm + s0;
    const int brev = (int)(__brev((unsigned)lane) >> 27);
    float vmax = 0.f;

    const float C1 = 0.30901699437494745f, C2 = -0.80901699437494745f;   // cos(2pi/5), cos(4pi/5)
    const float S1 = 0.95105651629515353f, S2 = 0.58778525229247314f;    // sin(2pi/5), sin(4pi/5)
    float* zre = sm.zre[warp];
    float* zim = sm.zim[warp];
    float* pw = sm.pw[warp];
    // filters own CONSECUTIVE bins: round A lane l reads bins binA + it, round B bins binB + 4*it
    const int binA = sm.mel_bin[lane], binB = sm.mel_bin[na * 32 + lane];

    for (int t = t_begin + warp; t < t_end; t += kWarpsA) {
        // ---- load + window: z[n] = x[2n] + i x[2n+1], n = 32*n1 + lane
        float2 v[5];
        const int64_t base = (int64_t)kHop * t - kHalf;
        if (base >= 0 && base + kNfft <= n_samples) {          // interior frame: no bounds checks
            const SampleT* __restrict__ xb = x + base + 2 * lane;
#pragma unroll
            for (int n1 = 0; n1 < 5; ++n1)
                v[n1] = make_float2(__fmul_rn((float)__ldg(xb + 64 * n1), hw[2 * n1]), __fmul_rn((float)__ldg(xb + 64 * n1 + 1), hw[2 * n1 + 1]));
        } else {
#pragma unroll
            for (int n1 = 0; n1 < 5; ++n1) {
                const int64_t i0 = base + 2 * (32 * n1 + lane), i1 = i0 + 1;
                const float a = (i0 >= 0 && i0 < n_samples) ? (float)__ldg(x + i0) : 0.f;
                const float b = (i1 >= 0 && i1 < n_samples) ? (float)__ldg(x + i1) : 0.f;
                v[n1] = make_float2(__fmul_rn(a, hw[2 * n1]), __fmul_rn(b, hw[2 * n1 + 1]));   // never contracted: f32 and s16 inputs agree bit for bit
            }
        }
        // ---- radix-5 over n1 (forward transform), then twiddle W_160^(lane*k1)
        float2 y[5];
        {
            float2 t1 = make_float2(v[1].x + v[4].x, v[1].y + v[4].y);
            float2 t2 = make_float2(v[2].x + v[3].x, v[2].y + v[3].y);
            float2 t3 = make_float2(v[1].x - v[4].x, v[1].y - v[4].y);
            float2 t4 = make_float2(v[2].x - v[3].x, v[2].y - v[3].y);
            y[0] = make_float2(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
            float2 m1 = make_float2(v[0].x + C1 * t1.x + C2 * t2.x, v[0].y + C1 * t1.y + C2 * t2.y);
            float2 m2 = make_float2(v[0].x + C2 * t1.x + C1 * t2.x, v[0].y + C2 * t1.y + C1 * t2.y);
            float2 q1 = make_float2(S1 * t3.x + S2 * t4.x, S1 * t3.y + S2 * t4.y);
            float2 q2 = make_float2(S2 * t3.x - S1 * t4.x, S2 * t3.y - S1 * t4.y);
            y[1] = make_float2(m1.x + q1.y, m1.y - q1.x);
            y[4] = make_float2(m1.x - q1.y, m1.y + q1.x);
            y[2] = make_float2(m2.x + q2.y, m2.y - q2.x);
            y[3] = make_float2(m2.x - q2.y, m2.y + q2.x);
        }
#pragma unroll
        for (int k1 = 1; k1 < 5; ++k1) y[k1] = cmul(y[k1], tw[k1]);
        // ---- 32-point DIF FFT across lanes for each k1 (output in bit-reversed lane order):
        //      every lane computes (sg*y + o) * w with its own sg, w -- no divergence
#pragma unroll
        for (int st = 0; st < 5; ++st) {
            const int h = 16 >> st;
#pragma unroll
            for (int k1 = 0; k1 < 5; ++k1) {
                const float ox = __shfl_xor_sync(FULL, y[k1].x, h);
                const float oy = __shfl_xor_sync(FULL, y[k1].y, h);
                const float2 a = make_float2(fmaf(sg[st], y[k1].x, ox), fmaf(sg[st], y[k1].y, oy));
                y[k1] = (st < 4) ? cmul(a, ws[st < 4 ? st : 0]) : a;
            }
        }
#pragma unroll
        for (int k1 = 0; k1 < 5; ++k1) { zre[k1 + 5 * brev] = y[k1].x; zim[k1 + 5 * brev] = y[k1].y; }
        __syncwarp();
        // ---- real-input post-pass + power spectrum: bins k = lane + 32*i, i < 5 (k = 0..159), then k = 160
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int k = lane + 32 * i;
            const int kb = (k == 0) ? 0 : kHalf - k;
            const float2 A = make_float2(zre[k], zim[k]);
            const float2 Bc = make_float2(zre[kb], zim[kb]);
            const float2 E = make_float2(0.5f * (A.x + Bc.x), 0.5f * (A.y - Bc.y));
            const float2 O = make_float2(0.5f * (A.y + Bc.y), -0.5f * (A.x - Bc.x));
            const float xr = E.x + O.x * wp[i].x - O.y * wp[i].y;
            const float xi = E.y + O.x * wp[i].y + O.y * wp[i].x;
            pw[k] = xr * xr + xi * xi;
        }
        if (lane == 0) { const float xn = zre[0] - zim[0]; pw[kHalf] = xn * xn; }   // Nyquist bin
        __syncwarp();
        // ---- mel filterbank (lane-balanced table)
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
        float* mo = mel_out + (f0 + t) * kMels;
        mo[lane] = accA;
        if ((lane & 3) == 0) mo[32 + (lane >> 2)] = accB;
        vmax = fmaxf(vmax, fmaxf(accA, accB));
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(FULL, vmax, o));
    if (lane == 0) atomicMax(reinterpret_cast<int*>(utt_max + u), __float_as_int(vmax));   // mel >= 0
}

// ------------------------------------------------------------------------------------------
// Kernel B
// ------------------------------------------------------------------------------------------
constexpr int kTileB = 64;
constexpr int kHaloB = kTileB + 8;

__global__ void __launch_bounds__(256)
mfcc_ceps_kernel(const float* __restrict__ mel, const float* __restrict__ utt_max,
                 const int64_t* __restrict__ frm_off, float* __restrict__ feat) {
    __shared__ float s_lm[kHaloB][kMels + 1];
    __shared__ float s_c[kHaloB][kCeps + 1];
    __shared__ float s_dct[kCeps][kMels + 1];
    __shared__ float s_out[kTileB * kFeat];
    const int u = blockIdx.x;
    const int64_t f0 = frm_off[u];
    const int T = (int)(frm_off[u + 1] - f0);
    const int t0 = blockIdx.y * kTileB;
    if (t0 >= T) return;
    const int t1 = min(T, t0 + kTileB);
    const int tid = threadIdx.x;
    const int g0 = max(0, min(t0 - 4, T - 9));
    const int g1 = min(T - 1, max(t1 + 3, 8));
    const int ng = g1 - g0 + 1;

    for (int i = tid; i < kCeps * kMels; i += blockDim.x) s_dct[i / kMels][i % kMels] = g_mfcc_tables.dct[i];
    const float ref_db = 10.0f * log10f(fmaxf(1e-10f, utt_max[u]));
    for (int i = tid; i < ng * kMels; i += blockDim.x) {
        const int j = i / kMels, m = i - j * kMels;
        const float v = mel[(f0 + g0 + j) * kMels + m];
        float db = 10.0f * log10f(fmaxf(1e-10f, v)) - ref_db;
        s_lm[j][m] = fmaxf(db, -80.0f);
    }
    __syncthreads();
    for (int i = tid; i < ng * kCeps; i += blockDim.x) {
        const int j = i / kCeps, k = i - j * kCeps;
        float acc = 0.f;
#pragma unroll 8
        for (int m = 0; m < kMels; ++m) acc = fmaf(s_dct[k][m], s_lm[j][m], acc);
        s_c[j][k] = acc;
    }
    __syncthreads();
    const int nt = t1 - t0;
    for (int i = tid; i < nt * kCeps; i += blockDim.x) {
        const int j = i / kCeps, k = i - j * kCeps;
        const int t = t0 + j;
        const int c = min(max(t, 4), T - 5) - g0;
        float d1 = 0.f, d2 = 0.f;
#pragma unroll
        for (int q = -4; q <= 4; ++q) {
            const float cv = s_c[c + q][k];
            d1 = fmaf((float)q * (1.0f / 60.0f), cv, d1);
            d2 = fmaf((float)(3 * q * q - 20) * (1.0f / 462.0f), cv, d2);
        }
        s_out[j * kFeat + kCeps + k] = d1;
        s_out[j * kFeat + 2 * kCeps + k] = d2;
    }
    for (int j = tid; j < nt; j += blockDim.x) {
        const int r = t0 + j - g0;
        float mean = 0.f;
#pragma unroll
        for (int k = 0; k < kCeps; ++k) mean += s_c[r][k];
        mean *= (1.0f / kCeps);
        float var = 0.f;
#pragma unroll
        for (int k = 0; k < kCeps; ++k) { const float d = s_c[r][k] - mean; var = fmaf(d, d, var); }
        const float inv = 1.0f / (sqrtf(var * (1.0f / kCeps)) + 1e-8f);
#pragma unroll
        for (int k = 0; k < kCeps; ++k) s_out[j * kFeat + k] = (s_c[r][k] - mean) * inv;
    }
    __syncthreads();
    float* __restrict__ dst = feat + (f0 + t0) * kFeat;
    for (int i = tid; i < nt * kFeat; i += blockDim.x) dst[i] = s_out[i];
}

}  // namespace loe

static int mfcc_launch(const void* pcm_dev, int pcm_format, const int64_t* pcm_off_dev, const int64_t* frm_off_dev,
                       int n_utt, int64_t total_frames, int max_frames, int min_frames,
                       const int32_t* mel_bin_dev, const float* mel_w_dev, int mel_na, int mel_nb,
                       float* mel_ws_dev, float* utt_max_dev, float* feat_dev, void* stream, int phases) {
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
    int st = ensure_tables();
    if (st != LOE_OK) return st;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (phases & 1) {
        LOE_CUDA(cudaMemsetAsync(utt_max_dev, 0, sizeof(float) * (size_t)n_utt, s));
        static_assert(sizeof(SmemA) <= 48 * 1024, "kernel A must fit the default dynamic shared memory limit");
        dim3 ga((unsigned)n_utt, (unsigned)((max_frames + kFramesPerBlockA - 1) / kFramesPerBlockA));
        const bool k16 = (mel_na == 11 && mel_nb == 5);      // the 16 kHz table of every reference call site
#define LOE_MEL_LAUNCH(T, A, B)                                                                                          \
        mfcc_mel_kernel<T, A, B><<<ga, kWarpsA * 32, sizeof(SmemA), s>>>((const T*)pcm_dev, pcm_off_dev, frm_off_dev, mel_bin_dev, \
                                                                        mel_w_dev, mel_na, mel_nb, mel_ws_dev, utt_max_dev)
        if (pcm_format == LOE_PCM_F32) { if (k16) LOE_MEL_LAUNCH(float, 11, 5); else LOE_MEL_LAUNCH(float, 0, 0); }
        else if (pcm_format == LOE_PCM_S16) { if (k16) LOE_MEL_LAUNCH(short, 11, 5); else LOE_MEL_LAUNCH(short, 0, 0); }
        else { set_error("unknown pcm_format %d", pcm_format); return LOE_ERR_VALUE; }
#undef LOE_MEL_LAUNCH
        LOE_LAUNCH_CHECK("mfcc_mel_kernel");
    }
    if (phases & 2) {
        dim3 gb((unsigned)n_utt, (unsigned)((max_frames + kTileB - 1) / kTileB));
        mfcc_ceps_kernel<<<gb, 256, 0, s>>>(mel_ws_dev, utt_max_dev, frm_off_dev, feat_dev);
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

extern "C" int loe_mfcc_phase_dev(const void* pcm_dev, int pcm_format, const int64_t* pcm_off_dev, const int64_t* frm_off_dev,
                                  int n_utt, int64_t total_frames, int max_frames, int min_frames,
                                  const int32_t* mel_bin_dev, const float* mel_w_dev, int mel_na, int mel_nb,
                                  float* mel_ws_dev, float* utt_max_dev, float* feat_dev, void* stream, int phases) {
    return mfcc_launch(pcm_dev, pcm_format, pcm_off_dev, frm_off_dev, n_utt, total_frames, max_frames, min_frames, mel_bin_dev,
                       mel_w_dev, mel_na, mel_nb, mel_ws_dev, utt_max_dev, feat_dev, stream, phases);
}
