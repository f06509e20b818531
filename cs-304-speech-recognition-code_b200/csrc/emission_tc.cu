// Gaussian emission scoring on the 5th-generation tensor cores (tcgen05 + TMEM), 3xTF32.
//
// Replaces MultivariateNormal.log_pdf (hidden_markov_model.py:46-48 -> scipy _logpdf) for a whole
// batch, recast as a dense contraction:
//
//     y[f, (s, j)] = sum_k  [x_f, 1][k] * W_s[k, j]          W_s = [ U_s ; -mu_s U_s ]   (K = 40)
//     score[f, s]  = cst_s - 0.5 * sum_j y[f, (s, j)]^2
//
// Precision: plain TF32 misses the 1e-4 bar by two orders of magnitude, so both operands are
// split x = x_hi + x_lo (x_hi = x rounded to TF32) and three MMAs hi*hi + lo*hi + hi*lo are
// accumulated in fp32 in TMEM (error ~3e-6 of the operand scale, measured).  Flops are counted
// once (2*40*39*S per frame) although three MMAs are issued.
//
// Decomposition: B-stationary.  CTA (n, g) owns one tile of 6 states (N = 240 columns; W hi/lo
// = 75 KB resident in shared memory for the whole kernel, loaded once) and walks the frame
// tiles g, g+G, ... (M = 128 frames).  The CTAs of the different state tiles walk the same frames
// at about the same pace, so a feature tile is fetched from HBM once and re-read from L2.  The SMs
// are divided over the state tiles in proportion to their MMA cost: the last tile may hold fewer
// states, is issued with a narrower N and gets fewer CTAs, so that every SM is busy to the end.
//   warps 0-7  producers: the raw [128 x 39] feature tile (19 968 contiguous bytes) arrives by
//              cp.async.bulk (TMA engine, mbarrier complete_tx), two tiles in flight; the warps
//              split it hi/lo and store it in the canonical K-major no-swizzle UMMA layout
//              (8 x 16 B core matrices; LBO = K-chunk stride, SBO = 128 B), fence.proxy.async,
//              arrive on a_full[stage]
//   warp  12   one thread issues 15 tcgen05.mma.kind::tf32 (M128 N240 K8) per tile and
//              tcgen05.commit's to a_empty[stage] and tmem_full[buf]
//   warps 8-11 epilogue: tcgen05.ld the 128 x 240 fp32 accumulator (thread = frame row),
//              square + sum each group of 40 columns, store 6 scores per frame
// Accumulators are double buffered in TMEM (2 x 240 of the 512 columns), A in shared memory
// (2 stages), so staging(i+1), MMA(i) and epilogue(i-1) overlap.
#include "tcgen05.cuh"

namespace loe {
namespace tc {

constexpr int kTileM = 128;
constexpr int kK = 40;
constexpr int kKChunks = kK / 4;        // 16-byte chunks along K
constexpr int kKSteps = kK / 8;         // one tf32 MMA consumes K = 8
constexpr int kColsPerState = 40;
constexpr int kStatesPerTile = 6;
constexpr int kTileN = kStatesPerTile * kColsPerState;   // 240
constexpr int kTmemCols = 512;
constexpr int kBufStride = 256;         // TMEM column offset between the two accumulators
constexpr int kDim = 39;
constexpr int kProducerThreads = 256;      // 8 warps: 2 per SM sub-partition, so staging latencies overlap
constexpr int kEpilogueThreads = 128;
constexpr int kThreads = kProducerThreads + kEpilogueThreads + 32;

constexpr int kALbo = kTileM * 16;      // 2048: byte stride between K-adjacent core matrices of A
constexpr int kBLbo = kTileN * 16;      // 3840
constexpr int kSbo = 128;               // byte stride between 8-row groups
constexpr int kABytes = kKChunks * kALbo;   // 20480
constexpr int kBBytes = kKChunks * kBLbo;   // 38400

struct __align__(128) Smem {
    uint8_t b_hi[kBBytes];
    uint8_t b_lo[kBBytes];
    uint8_t a_hi[2][kABytes];
    uint8_t a_lo[2][kABytes];
    float raw[2][kTileM * kDim];       // raw feature tiles, filled by cp.async.bulk (TMA engine), 2 in flight
    float cst[8];
    uint64_t raw_full[2], a_full[2], a_empty[2], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ float tf32_round(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

__global__ void __launch_bounds__(kThreads, 1)
emission_tc_kernel(const float* __restrict__ feat, int64_t n_frames, const float* __restrict__ b_packed,
                   const float* __restrict__ cst_pad, int n_states, float* __restrict__ out, int ld_out, int use_bulk,
                   int g_full, int g_last) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // CTA -> (state tile, frame-tile group).  Full tiles get g_full CTAs each, the last (possibly narrower,
    // hence cheaper) tile gets g_last, so that all SMs finish together.
    const int n_tiles = (n_states + kStatesPerTile - 1) / kStatesPerTile;
    const int cta = blockIdx.x;
    const int n_tile = min(cta / g_full, n_tiles - 1);
    const int G = (n_tile == n_tiles - 1) ? g_last : g_full;
    const int g = cta - n_tile * g_full;
    const int valid = min(kStatesPerTile, n_states - n_tile * kStatesPerTile);       // states of this tile
    const int n_cols = ((valid * kColsPerState + 15) / 16) * 16;                      // MMA N: multiple of 16
    const int n_mtiles = (int)((n_frames + kTileM - 1) / kTileM);

    // ---- one-time setup: barriers, TMEM, resident B tile
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
        const float4* src = reinterpret_cast<const float4*>(b_packed + (size_t)n_tile * (2 * kBBytes / 4));
        float4* dst = reinterpret_cast<float4*>(sm.b_hi);             // b_hi and b_lo are contiguous
        for (int i = tid; i < 2 * kBBytes / 16; i += kThreads) dst[i] = __ldg(src + i);
        if (tid < kStatesPerTile) sm.cst[tid] = cst_pad[n_tile * kStatesPerTile + tid];
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    if (warp < kProducerThreads / 32) {
        // =========================== producers ===========================
        constexpr int kTileElems = kTileM * kDim;                       // 4992 floats = 19968 B, a multiple of 16
        constexpr uint32_t kTileBytes = kTileElems * sizeof(float);
        const int n_it = (g < n_mtiles) ? (n_mtiles - g + G - 1) / G : 0;
        // bulk copies need a 16-byte aligned source (use_bulk) and a whole tile
        auto tile_full = [&](int it) { return use_bulk && (int64_t)(g + it * G + 1) * kTileM <= n_frames; };
        auto issue = [&](int it) {                       // thread 0 only: full tiles are contiguous and 16-byte aligned
            bulk_load(sm.raw[it & 1], feat + (int64_t)(g + it * G) * kTileElems, kTileBytes, &sm.raw_full[it & 1]);
        };
        if (tid == 0) {
            if (n_it > 0 && tile_full(0)) issue(0);
            if (n_it > 1 && tile_full(1)) issue(1);
        }
        const int row_id = tid & (kTileM - 1);
        const int kc0 = (tid >> 7) * (kKChunks / 2);       // this thread's half of the K chunks
        for (int it = 0; it < n_it; ++it) {
            const int s = it & 1;
            const uint32_t k = (uint32_t)(it >> 1);
            if (tile_full(it)) {
                mbar_wait(&sm.raw_full[s], k & 1);         // TMA bytes have landed
            } else {
                // the batch's last, partial tile: plain loads, rows beyond the end read as zero
                const int64_t f0 = (int64_t)(g + it * G) * kTileM;
                const int total = (int)(n_frames - f0) * kDim;
                for (int e = tid; e < kTileElems; e += kProducerThreads) sm.raw[s][e] = (e < total) ? __ldg(feat + f0 * kDim + e) : 0.f;
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            mbar_wait(&sm.a_empty[s], (k & 1) ^ 1);        // MMA finished reading this stage
            const float* row = sm.raw[s] + row_id * kDim;  // stride 39 words: conflict free
            uint8_t* ah = sm.a_hi[s] + row_id * 16;
            uint8_t* al = sm.a_lo[s] + row_id * 16;
#pragma unroll
            for (int kk = 0; kk < kKChunks / 2; ++kk) {
                const int kc = kc0 + kk;
                float v[4], h[4], l[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = kc * 4 + q;
                    v[q] = (c < kDim) ? row[c] : 1.0f;     // column 39: the constant 1 of the bias row
                    h[q] = tf32_round(v[q]);
                    l[q] = v[q] - h[q];
                }
                *reinterpret_cast<float4*>(ah + kc * kALbo) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4*>(al + kc * kALbo) = make_float4(l[0], l[1], l[2], l[3]);
            }
            fence_proxy_async();
            mbar_arrive(&sm.a_full[s]);
            asm volatile("bar.sync 1, 256;" ::: "memory"); // everyone done with raw[s]: it may be refilled
            if (tid == 0 && it + 2 < n_it && tile_full(it + 2)) issue(it + 2);
        }
    } else if (warp == (kProducerThreads + kEpilogueThreads) / 32) {
        // =========================== MMA issuer ===========================
        // whole warp in step, instructions predicated on one elected lane inside the asm: operands in uniform registers
        // (see emission_h16.cu)
        {
            const uint32_t leader = elect_one();
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
            const uint32_t b_hi = smem_u32(sm.b_hi), b_lo = smem_u32(sm.b_lo);
            int it = 0;
            for (int m = g; m < n_mtiles; m += G, ++it) {
                const int s = it & 1;
                const uint32_t k = (uint32_t)(it >> 1);
                mbar_wait(&sm.a_full[s], k & 1);
                mbar_wait(&sm.tmem_empty[s], (k & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(s * kBufStride);
                const uint32_t a_hi = smem_u32(sm.a_hi[s]), a_lo = smem_u32(sm.a_lo[s]);
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a = (pass == 1) ? a_lo : a_hi;
                    const uint32_t b = (pass == 2) ? b_lo : b_hi;
#pragma unroll
                    for (int ks = 0; ks < kKSteps; ++ks)
                        mma_tf32_elected(leader, d, make_desc(a + ks * 2 * kALbo, kALbo), make_desc(b + ks * 2 * kBLbo, kBLbo), idesc,
                                 (pass | ks) ? 1u : 0u);
                }
                mma_commit_elected(leader, &sm.a_empty[s]);
                mma_commit_elected(leader, &sm.tmem_full[s]);
            }
        }
    } else {
        // =========================== epilogue ===========================
        const int q = warp & 3;                             // TMEM lane quarter this warp may touch
        const int r = q * 32 + lane;
        int it = 0;
        for (int m = g; m < n_mtiles; m += G, ++it) {
            const int s = it & 1;
            const uint32_t k = (uint32_t)(it >> 1);
            mbar_wait(&sm.tmem_full[s], k & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * kBufStride);
            float score[kStatesPerTile];
#pragma unroll
            for (int j = 0; j < kStatesPerTile; ++j) {
                if (j >= valid) { score[j] = 0.f; continue; }      // warp-uniform: the narrow last tile reads less
                float v[40];
                tmem_ld32(taddr + j * kColsPerState, v);
                tmem_ld8(taddr + j * kColsPerState + 32, v + 32);
                tmem_ld_wait();
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < kDim; ++c) acc = fmaf(v[c], v[c], acc);
                score[j] = sm.cst[j] - 0.5f * acc;
            }
            tc_fence_before();
            mbar_arrive(&sm.tmem_empty[s]);
            const int64_t f = (int64_t)m * kTileM + r;
            if (f < n_frames) {
                float* o = out + f * ld_out + n_tile * kStatesPerTile;
#pragma unroll
                for (int j = 0; j < kStatesPerTile; ++j)
                    if (n_tile * kStatesPerTile + j < n_states) o[j] = score[j];
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

}  // namespace tc
}  // namespace loe

extern "C" int loe_emission_tc_tiles(int n_states) { return (n_states + loe::tc::kStatesPerTile - 1) / loe::tc::kStatesPerTile; }

extern "C" int loe_emission_tc_dev(const float* feat_dev, int64_t n_frames, int dim, const float* b_packed_dev,
                                   const float* cst_pad_dev, int n_states, float* out_dev, int ld_out, void* stream) {
    using namespace loe;
    using namespace loe::tc;
    if (n_frames <= 0 || n_states <= 0) return LOE_OK;
    if (dim != kDim) { set_error("tensor-core emission path is built for dim == 39 (got %d)", dim); return LOE_ERR_UNSUPPORTED; }
    if (ld_out < n_states) { set_error("ld_out (%d) < n_states (%d)", ld_out, n_states); return LOE_ERR_VALUE; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    static int sm_count[64] = {0};
    int dev = 0;
    LOE_CUDA(cudaGetDevice(&dev));
    if (dev >= 64) dev = 63;
    if (sm_count[dev] == 0) {
        LOE_CUDA(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
        LOE_CUDA(cudaFuncSetAttribute(emission_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    }
    int g_full = 1, g_last = 1;
    split_sms(n_states, n_frames, sm_count[dev], &g_full, &g_last);
    const int n_tiles = (n_states + kStatesPerTile - 1) / kStatesPerTile;
    const unsigned grid = (unsigned)((n_tiles - 1) * g_full + g_last);
    const int use_bulk = (reinterpret_cast<uintptr_t>(feat_dev) & 15) == 0 ? 1 : 0;
    emission_tc_kernel<<<grid, kThreads, sizeof(Smem), s>>>(feat_dev, n_frames, b_packed_dev, cst_pad_dev, n_states, out_dev, ld_out,
                                                            use_bulk, g_full, g_last);
    LOE_LAUNCH_CHECK("emission_tc_kernel");
    return LOE_OK;
}
