// Diagonal-covariance Gaussian-mixture emission scoring (BASELINE.json north_star kernel (2), configs[0] extension set and
// configs[4]: ~40 phones x 3 states, 16 Gaussians per state).
//
//     score[f, s] = log sum_m  w_sm N(x_f; mu_sm, diag sigma^2_sm)
//                 = LSE_m ( c_sm - 1/2 sum_k (x_fk - mu_smk)^2 / sigma^2_smk ),   c_sm = log w_sm - 1/2 (D log 2pi + sum_k log sigma^2_smk)
//
// The live reference scores full-covariance single Gaussians only (hidden_markov_model.py:20-48; emission*.cu); its
// abandoned GMM code gives the semantics -- logaddexp over log w_m + log N_m (deprecated/gaussian_mixture_model.py:157-162).
// Parity is against the restated NumPy oracle (oracle/gmm.py): parity unpinned by construction.
//
// Tensor-core path (loe_emission_gmm_tc_dev): the quadratic form is recast as ONE dense contraction
//     y[f, n] = [z^2 (39), 1, z (39), 0] . [ -t^2 / (2 sigma^2) ; c ; t mu' / sigma^2 ; 0 ][:, n]        K = 80, n = (state, mixture)
// with z = (x - shift) / t: per column tile a shift vector (the mean of the tile's component means: the products
// z^2 P and z Q cancel to the small quadratic form, the shift keeps them small) and per-dimension power-of-two scales
// t (so that a typical sigma / t is about 1 and the binary16 operands are used where they are dense).  Operands are
// split into two binary16 parts (22 significant bits) and hi*hi + lo*hi + hi*lo run as tcgen05.mma.kind::f16 with fp32
// accumulation in TMEM: 15 MMAs of M128 N240 K16 per 128-frame tile.  The epilogue thread (= frame row) reads the
// accumulator columns and does the log-sum-exp over each state's mixtures in registers.
// B-stationary CTAs: one column tile of 240 (state, mixture) columns resident in shared memory (77 KB), frame tiles
// streamed through it: raw feature tiles by cp.async.bulk (TMA engine), 4 producer warps (thread = row) build the A operand,
// one thread issues the MMAs, 12 epilogue warps in three warpgroups drain TMEM -- each warpgroup 80 columns = whole states
// of every row, so the exp / log work of a tile (30 720 exponentials at 16 mixtures) is spread over 12 warps instead of
// the 4 the TMEM lane rule suggests (the first version: epilogue-bound at 4 % of the tensor peak, ncu: profiles/).
// Rows whose |z| reaches 128 (z^2 near the binary16 range) or is not finite are computed by their producer thread with
// plain float32 arithmetic from the unpacked model: any input gives the SIMT kernel's answer.
//
// SIMT path (loe_emission_gmm_dev): float32 / float64, thread per (frame, state); the float64 instance is the exact mode.
#include <cuda_fp16.h>
#include "tcgen05.cuh"

namespace loe {
namespace gmm {
using namespace loe::tc;

constexpr int kDim = 39;
constexpr int kTileM = 128;
constexpr int kK = 80;                      // z^2 (39), 1, z (39), 0
constexpr int kChunksPerPart = kK / 8;      // 10 chunks of 8 halfs
constexpr int kAChunks = 2 * kChunksPerPart;    // hi 0-9, lo 10-19
constexpr int kBChunks = 2 * kChunksPerPart;
constexpr int kTileN = 240;                 // (state, mixture) columns per tile; mixtures padded to a power of two <= 16
constexpr int kTmemCols = 512;
constexpr int kBufStride = 256;
constexpr int kProducerThreads = 128;            // thread = frame row
constexpr int kEpiGroups = 3;                    // epilogue warpgroups: each drains 80 of the 240 accumulator columns of every row
constexpr int kEpiCols = kTileN / kEpiGroups;    // 80: a multiple of every padded mixture count (1, 2, 4, 8, 16)
constexpr int kEpilogueThreads = 128 * kEpiGroups;
constexpr int kThreads = kProducerThreads + kEpilogueThreads + 32;   // + the MMA warp.  17 warps are allocated as 20 (groups of four):
                                                                     // 96 registers per thread, which a 48-column drain round fits
constexpr int kALbo = kTileM * 16;          // 2048
constexpr int kBLbo = kTileN * 16;          // 3840
constexpr int kABytes = kAChunks * kALbo;   // 40960
constexpr int kBBytes = kBChunks * kBLbo;   // 76800
constexpr float kZMax = 128.0f;

struct __align__(128) Smem {
    uint8_t b[kBBytes];
    uint8_t a[2][kABytes];
    float raw[2][kTileM * kDim];
    float shift[40], iscale[40];
    uint8_t slow[4][kTileM];                // rows the tensor path cannot take (slot = tile & 3)
    uint64_t raw_full[2], a_full[2], a_empty[2], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};
static_assert(sizeof(Smem) <= 227 * 1024, "shared memory of the GMM emission kernel");

// hi / lo binary16 split of 8 consecutive values into chunk kc of the hi part and of the lo part of one A row
__device__ __forceinline__ void split_store(const float* x, uint8_t* a_row, int kc) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __half2 hh = __floats2half2_rn(x[2 * q], x[2 * q + 1]);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(x[2 * q] - hf.x, x[2 * q + 1] - hf.y);
        h[q] = *reinterpret_cast<const uint32_t*>(&hh);
        l[q] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(a_row + kc * kALbo) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(a_row + (kChunksPerPart + kc) * kALbo) = make_uint4(l[0], l[1], l[2], l[3]);
}

// log sum exp over the first n_mix of MP consecutive accumulator columns, branch-free: 4 instructions per component (FMNMX,
// FFMA, MUFU.EX2, FADD).  FULL: n_mix == MP, no padding columns; otherwise a padding column reads as -inf.  The exponent
// is formed as fma(v, log2 e, -m log2 e): the rounding of m log2 e scales every term alike (2^-24 |m| on the result).
// All components -inf gives -inf (the clamp keeps inf - inf out), a NaN component gives NaN.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned long long pk2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk2(unsigned long long p, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
template <int MP, bool FULL>
__device__ __forceinline__ float lse(const float* v, int n_mix) {
    if (MP == 1) return v[0];
    constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
    float vv[MP];
#pragma unroll
    for (int i = 0; i < MP; ++i) vv[i] = (FULL || i < n_mix) ? v[i] : -CUDART_INF_F;
    float m = vv[0];
#pragma unroll
    for (int i = 1; i < MP; ++i) m = fmaxf(m, vv[i]);
    m = fmaxf(m, -3.0e38f);
    const float nms = -m * kLog2e;
    // accumulator columns come in aligned register pairs: the scaling and the summation run as packed f32x2 instructions
    const unsigned long long l2 = pk2(kLog2e, kLog2e), n2 = pk2(nms, nms);
    unsigned long long acc[2] = {0ull, 0ull};
#pragma unroll
    for (int i = 0; i < MP; i += 2) {
        float t0, t1;
        upk2(fma2(pk2(vv[i], vv[i + 1]), l2, n2), t0, t1);
        acc[(i >> 1) & 1] = add2(acc[(i >> 1) & 1], pk2(ex2_approx(t0), ex2_approx(t1)));
    }
    float s0, s1;
    upk2(MP > 2 ? add2(acc[0], acc[1]) : acc[0], s0, s1);
    return fmaf(lg2_approx(s0 + s1), kLn2, m);
}

// exact float32 evaluation of one (frame, state): the slow path of the tensor-core kernel and the body of the SIMT kernel
template <typename T>
__device__ __forceinline__ float score_state(const T* x, const T* __restrict__ mean, const T* __restrict__ inv_var,
                                             const T* __restrict__ cst, int s, int n_mix) {
    T best = -CUDART_INF;
    T vals[16];
    bool nan_seen = false;
    for (int m = 0; m < n_mix; ++m) {
        const T* mu = mean + (size_t)(s * n_mix + m) * kDim;
        const T* iv = inv_var + (size_t)(s * n_mix + m) * kDim;
        T q = 0;
#pragma unroll
        for (int k = 0; k < kDim; ++k) { const T d = x[k] - mu[k]; q = fma(d * d, iv[k], q); }
        vals[m] = cst[s * n_mix + m] - (T)0.5 * q;
        nan_seen |= vals[m] != vals[m];
        best = vals[m] > best ? vals[m] : best;
    }
    if (nan_seen) return CUDART_NAN_F;                          // NaN in, NaN out (like the oracle's logsumexp)
    if (!(best > -CUDART_INF)) return (float)best;              // all components -inf: no exp of inf - inf
    T sum = 0;
    for (int m = 0; m < n_mix; ++m) sum += exp(vals[m] - best);
    return (float)(best + log(sum));
}

// the float32 path of one out-of-range row (rare): kept out of line so that its registers (the feature row, the component
// values) do not weigh on the epilogue loop
__device__ __noinline__ void score_row_slow(const float* __restrict__ feat_row, const float* __restrict__ mean32, const float* __restrict__ inv_var32,
                                            const float* __restrict__ cst32, int s_begin, int s_end, int n_mix, float* __restrict__ o) {
    float x[kDim];
#pragma unroll
    for (int c = 0; c < kDim; ++c) x[c] = __ldg(feat_row + c);
    for (int s = s_begin; s < s_end; ++s) o[s - s_begin] = score_state<float>(x, mean32, inv_var32, cst32, s, n_mix);
}

template <int MP, bool FULL>
__global__ void __launch_bounds__(kThreads, 1)
emission_gmm_tc_kernel(const float* __restrict__ feat, int64_t n_frames, const uint8_t* __restrict__ b_packed,
                       const float* __restrict__ shift_scale, const float* __restrict__ mean32, const float* __restrict__ inv_var32,
                       const float* __restrict__ cst32, int n_states, int n_mix, float* __restrict__ out, int ld_out, int use_bulk,
                       int g_full, int g_last) {
    constexpr int SPT = kTileN / MP;                                 // states per column tile
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = (n_states + SPT - 1) / SPT;
    const int cta = blockIdx.x;
    const int n_tile = min(cta / g_full, n_tiles - 1);
    const int G = (n_tile == n_tiles - 1) ? g_last : g_full;
    const int g = cta - n_tile * g_full;
    const int valid = min(SPT, n_states - n_tile * SPT);             // states of this tile
    const int n_cols = ((valid * MP + 15) / 16) * 16;                // MMA N: multiple of 16
    const int n_mtiles = (int)((n_frames + kTileM - 1) / kTileM);

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.raw_full[i], 1);
            mbar_init(&sm.a_full[i], kProducerThreads);
            mbar_init(&sm.a_empty[i], 1);
            mbar_init(&sm.tmem_full[i], 1);
            mbar_init(&sm.tmem_empty[i], kEpilogueThreads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kProducerThreads / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    {
        const uint4* src = reinterpret_cast<const uint4*>(b_packed + (size_t)n_tile * kBBytes);
        uint4* dst = reinterpret_cast<uint4*>(sm.b);
        for (int i = tid; i < kBBytes / 16; i += kThreads) dst[i] = __ldg(src + i);
        if (tid < 40) { sm.shift[tid] = shift_scale[n_tile * 80 + tid]; sm.iscale[tid] = shift_scale[n_tile * 80 + 40 + tid]; }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    if (warp < kProducerThreads / 32) {
        // =========================== producers ===========================
        constexpr int kTileElems = kTileM * kDim;
        constexpr uint32_t kTileBytes = kTileElems * sizeof(float);
        const int n_it = (g < n_mtiles) ? (n_mtiles - g + G - 1) / G : 0;
        auto tile_full = [&](int it) { return use_bulk && (int64_t)(g + it * G + 1) * kTileM <= n_frames; };
        auto issue = [&](int it) {
            bulk_load(sm.raw[it & 1], feat + (int64_t)(g + it * G) * kTileElems, kTileBytes, &sm.raw_full[it & 1]);
        };
        if (tid == 0) {
            if (n_it > 0 && tile_full(0)) issue(0);
            if (n_it > 1 && tile_full(1)) issue(1);
        }
        const int row_id = tid;
        for (int it = 0; it < n_it; ++it) {
            const int s = it & 1;
            const uint32_t k = (uint32_t)(it >> 1);
            if (tile_full(it)) {
                mbar_wait(&sm.raw_full[s], k & 1);
            } else {
                const int64_t f0 = (int64_t)(g + it * G) * kTileM;
                const int total = (int)(n_frames - f0) * kDim;
                for (int e = tid; e < kTileElems; e += kProducerThreads) sm.raw[s][e] = (e < total) ? __ldg(feat + f0 * kDim + e) : 0.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            mbar_wait(&sm.a_empty[s], (k & 1) ^ 1);
            const float* row = sm.raw[s] + row_id * kDim;
            uint8_t* a_row = sm.a[s] + row_id * 16;
            float mx = 0.f;
            // chunk by chunk (8 columns): z, then [z^2 | 1] into chunk kc and [z | 0] into chunk 5 + kc.  A row that turns out
            // to be out of range (or not finite) is flagged afterwards: its operand may hold inf / NaN, which stays inside
            // its own accumulator row, and that row is never read -- its epilogue threads score it in float32.
#pragma unroll
            for (int kc = 0; kc < 5; ++kc) {
                float z[8], z2[8];
#pragma unroll
                for (int q8 = 0; q8 < 8; ++q8) {
                    const int c = kc * 8 + q8;
                    if (c < kDim) {
                        z[q8] = (row[c] - sm.shift[c]) * sm.iscale[c];
                        z2[q8] = z[q8] * z[q8];
                        mx = fmaxf(mx, fabsf(z[q8]));
                    } else {
                        z[q8] = 0.f;                 // column 39: the constant 1 sits in the z^2 half
                        z2[q8] = 1.0f;
                    }
                }
                split_store(z2, a_row, kc);
                split_store(z, a_row, 5 + kc);
            }
            const bool slow = !(mx < kZMax);                            // NaN compares false: flagged
            sm.slow[it & 3][row_id] = slow ? 1 : 0;
            fence_proxy_async();
            mbar_arrive(&sm.a_full[s]);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tid == 0 && it + 2 < n_it && tile_full(it + 2)) issue(it + 2);
            // an out-of-range row is scored here, by the thread that found it, in float32 from the unpacked model (rare;
            // the producers have slack): the epilogue threads skip it
            const int64_t f = (int64_t)(g + it * G) * kTileM + row_id;
            if (slow && f < n_frames)
                score_row_slow(feat + f * kDim, mean32, inv_var32, cst32, n_tile * SPT, n_tile * SPT + valid, n_mix, out + f * ld_out + n_tile * SPT);
        }
    } else if (warp == (kProducerThreads + kEpilogueThreads) / 32) {
        // =========================== MMA issuer ===========================
        // whole warp in step, instructions predicated on one elected lane inside the asm: operands in uniform registers
        // (see emission_h16.cu)
        {
            const uint32_t leader = elect_one();
            // c_format F32 (bit 4), a/b format F16 (0), N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (1u << 4) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
            const uint32_t b_base = smem_u32(sm.b);
            int it = 0;
            for (int m = g; m < n_mtiles; m += G, ++it) {
                const int s = it & 1;
                const uint32_t k = (uint32_t)(it >> 1);
                mbar_wait(&sm.a_full[s], k & 1);
                mbar_wait(&sm.tmem_empty[s], (k & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(s * kBufStride);
                const uint32_t a_base = smem_u32(sm.a[s]);
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {                        // hi*hi, lo*hi, hi*lo
                    const uint32_t a = a_base + ((pass == 1) ? kChunksPerPart * kALbo : 0);
                    const uint32_t b = b_base + ((pass == 2) ? kChunksPerPart * kBLbo : 0);
#pragma unroll
                    for (int ks = 0; ks < kChunksPerPart / 2; ++ks)
                        mma_f16_elected(leader, d, make_desc(a + ks * 2 * kALbo, kALbo), make_desc(b + ks * 2 * kBLbo, kBLbo), idesc, (pass | ks) ? 1u : 0u);
                }
                mma_commit_elected(leader, &sm.a_empty[s]);
                mma_commit_elected(leader, &sm.tmem_full[s]);
            }
        }
    } else {
        // =========================== epilogue ===========================
        // Three warpgroups; warp w may touch the TMEM lanes 32 (w % 4) .. + 31 (its rows), warpgroup grp takes the columns
        // [80 grp, 80 grp + 80) of every row: whole states (80 is a multiple of MP), so every log-sum-exp stays inside one
        // thread and twelve warps share the exp / log work of a tile.
        const int q = warp & 3;
        const int grp = (warp - kProducerThreads / 32) >> 2;
        const int r = q * 32 + lane;
        const int c0 = grp * kEpiCols;
        const int st0 = c0 / MP;                                      // first state (local to the tile) of this warpgroup
        const int n_mine = min(max(valid - st0, 0), kEpiCols / MP);   // states of this warpgroup's slice that exist
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        float* o_col = out + n_tile * SPT + st0;
        int it = 0;
        for (int m = g; m < n_mtiles; m += G, ++it) {
            const int s = it & 1;
            mbar_wait(&sm.tmem_full[s], (uint32_t)(it >> 1) & 1);
            tc_fence_after();
            // two rounds (48 + 32 columns: both multiples of MP): the accumulator slice of a thread stays at 48 registers
            const uint32_t taddr = taddr0 + (uint32_t)(s * kBufStride);
            const int64_t f = (int64_t)m * kTileM + r;
            const bool work = f < n_frames && sm.slow[it & 3][r] == 0;
            float* o = o_col + (f < n_frames ? f : 0) * ld_out;
            constexpr int kR0 = 48, kS0 = kR0 / MP;                   // columns / states of the first round
            float v[kR0];
            if (n_mine > 0) {                                         // warp-uniform: a narrow last tile reads nothing here
                tmem_ld32(taddr, v);
                tmem_ld16(taddr + 32, v + 32);
                tmem_ld_wait();
            }
            if (work) {
#pragma unroll
                for (int j = 0; j < kS0; ++j)
                    if (j < n_mine) o[j] = lse<MP, FULL>(v + j * MP, n_mix);
            }
            if (n_mine > kS0) {
                tmem_ld32(taddr + kR0, v);
                tmem_ld_wait();
            }
            tc_fence_before();
            mbar_arrive(&sm.tmem_empty[s]);                           // the accumulator has been read: hand it back
            if (work) {
#pragma unroll
                for (int j = kS0; j < kEpiCols / MP; ++j)
                    if (j < n_mine) o[j] = lse<MP, FULL>(v + (j - kS0) * MP, n_mix);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kProducerThreads / 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// SIMT: thread per (frame, state)
template <typename T>
__global__ void __launch_bounds__(128)
emission_gmm_simt_kernel(const float* __restrict__ feat, int64_t n_frames, const T* __restrict__ mean, const T* __restrict__ inv_var,
                         const T* __restrict__ cst, int n_states, int n_mix, float* __restrict__ out, int ld_out) {
    const int64_t f = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (f >= n_frames) return;
    T x[kDim];
#pragma unroll
    for (int c = 0; c < kDim; ++c) x[c] = (T)__ldg(feat + f * kDim + c);
    for (int s = blockIdx.y; s < n_states; s += gridDim.y) out[f * ld_out + s] = score_state<T>(x, mean, inv_var, cst, s, n_mix);
}

static int padded_mix(int n_mix) {
    int mp = 1;
    while (mp < n_mix) mp <<= 1;
    return mp;
}

template <int MP, bool FULL>
static int launch_tc_impl(const float* feat_dev, int64_t n_frames, const void* b_packed_dev, const float* shift_scale_dev,
                     const float* mean32_dev, const float* inv_var32_dev, const float* cst32_dev, int n_states, int n_mix,
                     float* out_dev, int ld_out, cudaStream_t s) {
    constexpr int SPT = kTileN / MP;
    int dev = 0, sms = 0;
    LOE_CUDA(cudaGetDevice(&dev));
    LOE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    LOE_CUDA(cudaFuncSetAttribute(emission_gmm_tc_kernel<MP, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    const int n_tiles = (n_states + SPT - 1) / SPT;
    const int n_mtiles = (int)((n_frames + kTileM - 1) / kTileM);
    const int valid_last = n_states - (n_tiles - 1) * SPT;
    const double cost_last = (double)(((valid_last * MP + 15) / 16) * 16) / (double)kTileN;
    int g_full = 1, g_last = 1;
    if (n_tiles == 1) {
        g_last = sms;
    } else if (sms >= n_tiles) {
        double best = 1e30;
        for (int gf = 1; (n_tiles - 1) * gf < sms; ++gf) {
            const int gl = sms - (n_tiles - 1) * gf;
            const double t = (1.0 / gf > cost_last / gl) ? 1.0 / gf : cost_last / gl;
            if (t < best) { best = t; g_full = gf; g_last = gl; }
        }
    }
    if (g_full > n_mtiles) g_full = n_mtiles;
    if (g_last > n_mtiles) g_last = n_mtiles;
    const unsigned grid = (unsigned)((n_tiles - 1) * g_full + g_last);
    const int use_bulk = (reinterpret_cast<uintptr_t>(feat_dev) & 15) == 0 ? 1 : 0;
    emission_gmm_tc_kernel<MP, FULL><<<grid, kThreads, sizeof(Smem), s>>>(feat_dev, n_frames, static_cast<const uint8_t*>(b_packed_dev),
                                                                   shift_scale_dev, mean32_dev, inv_var32_dev, cst32_dev, n_states, n_mix,
                                                                   out_dev, ld_out, use_bulk, g_full, g_last);
    LOE_LAUNCH_CHECK("emission_gmm_tc_kernel");
    return LOE_OK;
}

template <int MP, typename... Args>
static int launch_tc(int n_mix, Args... args) {
    return n_mix == MP ? launch_tc_impl<MP, true>(args...) : launch_tc_impl<MP, false>(args...);
}

}  // namespace gmm
}  // namespace loe

extern "C" int loe_emission_gmm_tile_bytes(void) { return loe::gmm::kBBytes; }

extern "C" int loe_emission_gmm_tiles(int n_states, int n_mix) {
    if (n_states <= 0 || n_mix <= 0 || n_mix > 16) return 0;
    const int spt = loe::gmm::kTileN / loe::gmm::padded_mix(n_mix);
    return (n_states + spt - 1) / spt;
}

extern "C" int loe_emission_gmm_dev(const float* feat_dev, int64_t n_frames, int dim, const void* mean_dev, const void* inv_var_dev,
                                    const void* cst_dev, int n_states, int n_mix, float* out_dev, int ld_out, int precision, void* stream) {
    using namespace loe;
    using namespace loe::gmm;
    if (n_frames <= 0 || n_states <= 0) return LOE_OK;
    if (dim != kDim) { set_error("GMM emission kernels are built for dim == 39 (got %d)", dim); return LOE_ERR_UNSUPPORTED; }
    if (n_mix <= 0 || n_mix > 16) { set_error("1..16 mixtures per state supported (got %d)", n_mix); return LOE_ERR_UNSUPPORTED; }
    if (ld_out < n_states) { set_error("ld_out (%d) < n_states (%d)", ld_out, n_states); return LOE_ERR_VALUE; }
    if (precision != 0 && precision != 1) { set_error("unknown precision %d", precision); return LOE_ERR_VALUE; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    dim3 grid((unsigned)((n_frames + 127) / 128), (unsigned)(n_states < 8 ? n_states : 8));
    if (precision == 0)
        emission_gmm_simt_kernel<float><<<grid, 128, 0, s>>>(feat_dev, n_frames, (const float*)mean_dev, (const float*)inv_var_dev,
                                                             (const float*)cst_dev, n_states, n_mix, out_dev, ld_out);
    else
        emission_gmm_simt_kernel<double><<<grid, 128, 0, s>>>(feat_dev, n_frames, (const double*)mean_dev, (const double*)inv_var_dev,
                                                              (const double*)cst_dev, n_states, n_mix, out_dev, ld_out);
    LOE_LAUNCH_CHECK("emission_gmm_simt_kernel");
    return LOE_OK;
}

extern "C" int loe_emission_gmm_tc_dev(const float* feat_dev, int64_t n_frames, int dim, const void* b_packed_dev,
                                       const float* shift_scale_dev, const float* mean32_dev, const float* inv_var32_dev,
                                       const float* cst32_dev, int n_states, int n_mix, float* out_dev, int ld_out, void* stream) {
    using namespace loe;
    using namespace loe::gmm;
    if (n_frames <= 0 || n_states <= 0) return LOE_OK;
    if (dim != kDim) { set_error("GMM emission kernels are built for dim == 39 (got %d)", dim); return LOE_ERR_UNSUPPORTED; }
    if (n_mix <= 0 || n_mix > 16) { set_error("1..16 mixtures per state supported (got %d)", n_mix); return LOE_ERR_UNSUPPORTED; }
    if (ld_out < n_states) { set_error("ld_out (%d) < n_states (%d)", ld_out, n_states); return LOE_ERR_VALUE; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    switch (padded_mix(n_mix)) {
        case 1: return launch_tc<1>(n_mix, feat_dev, n_frames, b_packed_dev, shift_scale_dev, mean32_dev, inv_var32_dev, cst32_dev, n_states, n_mix, out_dev, ld_out, s);
        case 2: return launch_tc<2>(n_mix, feat_dev, n_frames, b_packed_dev, shift_scale_dev, mean32_dev, inv_var32_dev, cst32_dev, n_states, n_mix, out_dev, ld_out, s);
        case 4: return launch_tc<4>(n_mix, feat_dev, n_frames, b_packed_dev, shift_scale_dev, mean32_dev, inv_var32_dev, cst32_dev, n_states, n_mix, out_dev, ld_out, s);
        case 8: return launch_tc<8>(n_mix, feat_dev, n_frames, b_packed_dev, shift_scale_dev, mean32_dev, inv_var32_dev, cst32_dev, n_states, n_mix, out_dev, ld_out, s);
        default: return launch_tc<16>(n_mix, feat_dev, n_frames, b_packed_dev, shift_scale_dev, mean32_dev, inv_var32_dev, cst32_dev, n_states, n_mix, out_dev, ld_out, s);
    }
}
