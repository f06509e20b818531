// Gaussian emission scoring on the 5th-generation tensor cores, 3xFP16 variant of emission_tc.cu.
//
// Same contraction and the same split-operand idea as the 3xTF32 kernel
//     y[f, (s, j)] = sum_k [x_f, 1][k] * W_s[k, j],   score[f, s] = cst_s - 0.5 * sum_j y^2
// but the operands are split into two binary16 parts (x = hi + lo, 11 + 11 significant bits: the
// same 22 bits a TF32 pair carries) and the products hi*hi + lo*hi + hi*lo run as kind::f16 MMAs,
// which the tensor pipe executes at twice the TF32 rate.  The 15 (A chunk, B chunk) products of
// 8 halfs each are paired into 8 MMAs of K = 16 through the descriptors' leading byte offset
// (one zero A chunk and a second copy of the hi B chunks make the pairing come out even), against 15 TF32 MMAs.
//
// W_s is not scipy's U_s itself but the lower-triangular R_s^T of U_s^T = Q R_s (the host packs it; same
// |W^T (x - mean)|): with the accumulator columns ordered [column block][state][8] the MMAs of feature chunk c
// only run over the 48 (c + 1) columns that chunk can reach -- 60 % of the dense MMA work (see pair_chunk).
//
// Range: binary16 tops out at 65504.  The B image is checked on the host when it is packed (the
// caller falls back to the TF32 image otherwise); every feature row whose largest magnitude reaches
// 2^15 is scaled by an exact power of two before the split and the squared norm is scaled back in
// the epilogue, so the A operand cannot overflow.  Residuals below 2^-24 flush to the binary16
// subnormal grid: an absolute error of 3e-8 on values below 0.125, the same size as the TF32 pair's
// relative error there.
//
// Decomposition: B-stationary CTAs, each with the images of kHalves = 2 six-state column tiles resident; a staged
// 128-frame feature tile is multiplied with both (16 MMAs).  Warp roles: 8 producer warps in two groups (raw feature
// tiles by cp.async.bulk, one in flight per group; hi / lo split into the A stages), one MMA-issuing thread, 4 epilogue
// warps (two accumulators of 240 columns in TMEM, each drained in two 120-column rounds).  The register file is
// split with setmaxnreg: 88 per producer thread, 208 per epilogue thread -- the spill-free point; this kernel loses
// 10-30 % as soon as ptxas spills inside a role loop (build.py warns).
#include <cuda_fp16.h>
#include "tcgen05.cuh"
#include "h16_stage.cuh"

namespace loe {
namespace h16 {
using namespace loe::tc;

constexpr int kTileM = 128;
constexpr int kDim = 39;
constexpr int kK = 40;                  // 39 features + the constant 1 of the bias row
constexpr int kChunksPerPart = kK / 8;  // 16-byte chunks (8 halfs) of one hi or lo part: 5
constexpr int kAChunks = 2 * kChunksPerPart + 1;   // hi 0-4, lo 5-9, zero 10
constexpr int kBChunks = 3 * kChunksPerPart;       // hi 0-4, lo 5-9, second copy of hi at 10-14
constexpr int kColsPerState = 40;
constexpr int kStatesPerTile = 6;
constexpr int kTileN = kStatesPerTile * kColsPerState;   // 240
constexpr int kTmemCols = 512;
constexpr int kBufStride = 256;
constexpr int kProducerGroups = 2;           // groups of 128 threads (thread = feature row) taking the tiles in turn
constexpr int kProducerThreads = kProducerGroups * 128;
constexpr int kEpilogueThreads = 128;
constexpr int kThreads = kProducerThreads + kEpilogueThreads + 32;
constexpr int kALbo = kTileM * 16;      // 2048 B between K-adjacent A chunks
constexpr int kBLbo = kTileN * 16;      // 3840 B between K-adjacent B chunks
constexpr int kABytes = kAChunks * kALbo;   // 22528
constexpr int kBBytes = kBChunks * kBLbo;   // 57600
constexpr int kProducerRegs = 88;
constexpr int kEpilogueRegs = 208;
constexpr int kNumMma = 8;
static_assert(kALbo == kStageLbo && kChunksPerPart == kStageChunksPerPart && kK == kStageK, "h16_stage.cuh describes this kernel's A operand");
#ifndef LOE_H16_RAW_BUFS
#define LOE_H16_RAW_BUFS 2
#endif
constexpr int kRawBufs = LOE_H16_RAW_BUFS;     // a power of two, a multiple of the producer groups
static_assert((kRawBufs & (kRawBufs - 1)) == 0 && kRawBufs >= 2, "raw buffer ring");
#ifndef LOE_H16_HALVES
#define LOE_H16_HALVES 2
#endif
constexpr int kHalves = LOE_H16_HALVES;              // 6-state column tiles per CTA: a staged feature tile is multiplied with both

struct __align__(128) Smem {
    uint8_t b[kHalves][kBBytes];
    uint8_t a[2][kABytes];
    float raw[kRawBufs][kTileM * kDim];  // raw feature tiles, filled by cp.async.bulk: two in flight per producer group
    float inv2[4][kTileM];              // 4^e of the rows scaled by 2^-e (1 for ordinary rows); slot = tile & 3: the
                                        // producers run at most 3 tiles ahead of the epilogue
    float out_stage[4][32 * kStatesPerTile];   // per epilogue warp: 32 rows x 6 scores on their way to global memory
    float cst[kHalves * kStatesPerTile];
    uint64_t raw_full[kRawBufs], a_full[2], a_empty[2], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

// W_s is lower triangular (feature k feeds the columns j <= k; the bias row 39 feeds all), and the accumulator
// columns are ordered in five blocks of 48, [column block b = j / 8][state][j % 8]: the K chunk c (features 8c ..
// 8c + 7) only reaches the blocks 0 .. c, so its MMAs are issued with N = 48 (c + 1) instead of 240 -- 60 % of the
// dense MMA work.  The 15 (A chunk, B chunk) products pair up into 8 MMAs of K = 16:
//   (first A chunk, second A chunk, first B chunk, second B chunk, column blocks), widest first -- the first MMA
//   overwrites all 240 columns, the others accumulate.  An MMA runs at the width of its wider product, so products of
//   (nearly) equal width share one: 48 half-MMA column blocks are issued for the 45 the products need (the pairing of the
//   first version of this kernel issued 52: 1.294 -> 1.270 ms per 3.84 M frames).
constexpr int kBlockCols = kStatesPerTile * 8;      // 48 accumulator columns per column block
__host__ __device__ constexpr int pair_chunk(int mma, int which) {
    constexpr int t[kNumMma][5] = {
        {4, 9, 4, 14, 5},          // hi4 * hi4 + lo4 * hi4
        {3, 4, 8, 9, 5},           // hi3 * lo3 + hi4 * lo4
        {3, 8, 3, 13, 4},          // hi3 * hi3 + lo3 * hi3
        {2, 7, 2, 12, 3},          // hi2 * hi2 + lo2 * hi2
        {1, 2, 6, 7, 3},           // hi1 * lo1 + hi2 * lo2
        {1, 6, 1, 11, 2},          // hi1 * hi1 + lo1 * hi1
        {0, 5, 0, 10, 1},          // hi0 * hi0 + lo0 * hi0
        {0, 10, 5, 14, 1},         // hi0 * lo0 + zero * (anything finite)
    };
    return t[mma][which];
}
// state that owns accumulator column n (a pair of neighbouring columns never straddles two states)
__host__ __device__ constexpr int state_of(int n) { return (n % kBlockCols) / 8; }
static_assert(kBlockCols % 16 == 0 && 5 * kBlockCols == kTileN, "MMA N granularity / blocks cover the tile");

__device__ __forceinline__ unsigned long long pack_f2(float x, float y) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ float2 unpack_f2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
// acc += x * x on both fp32 lanes of a 64-bit register pair (FFMA2)
__device__ __forceinline__ unsigned long long ffma2(unsigned long long x, unsigned long long acc) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(r) : "l"(x), "l"(acc));
    return r;
}

// sum of squares of the first 39 of 40 accumulator columns
__device__ __forceinline__ float sumsq39(const float* v) {
    // packed fp32x2 FMAs (two lanes per instruction), two independent chains of pairs
    unsigned long long a = 0ull, b = 0ull;
#pragma unroll
    for (int c = 0; c < 36; c += 4) {
        a = ffma2(pack_f2(v[c], v[c + 1]), a);
        b = ffma2(pack_f2(v[c + 2], v[c + 3]), b);
    }
    a = ffma2(pack_f2(v[36], v[37]), a);
    const float2 fa = unpack_f2(a), fb = unpack_f2(b);
    return fmaf(v[38], v[38], (fa.x + fa.y) + (fb.x + fb.y));
}

// The CTA is launched with 128 registers per thread (13 warps are allocated like 16: 16 x 32 x 128 is the whole
// file).  The two producer warpgroups hand 40 registers per thread back (setmaxnreg.dec) and the epilogue
// warpgroup takes them (setmaxnreg.inc) for its 120-column accumulator slice.
// IMG = true: ``feat`` is not the feature matrix but the pre-split A image the cepstrum kernel wrote (h16_stage.cuh: per
// 128-frame tile 20 480 bytes = hi chunks 0-4, lo chunks 5-9 in the layout of an A stage) and ``inv2_g`` the row scales: the
// producer warps have nothing to do but ONE thread that bulk-copies tile after tile straight into the A stages.
template <bool IMG>
__device__ __forceinline__ void
emission_h16_body(const float* __restrict__ feat, const float* __restrict__ inv2_g, int64_t n_frames, const uint8_t* __restrict__ b_packed,
                  const float* __restrict__ cst_pad, int n_states, float* __restrict__ out, int ld_out, int use_bulk,
                  int g_full, int g_last, int cta) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // CTA -> (column supertile = kHalves state tiles of 6, frame-tile group).  Full supertiles get g_full CTAs each,
    // the last (possibly narrower, hence cheaper) one gets g_last, so that all SMs finish together.
    const int n_tiles = (n_states + kStatesPerTile - 1) / kStatesPerTile;
    const int n_super = (n_tiles + kHalves - 1) / kHalves;
    const int sup = min(cta / g_full, n_super - 1);
    const int G = (sup == n_super - 1) ? g_last : g_full;
    const int g = cta - sup * g_full;
    const int H = min(kHalves, n_tiles - sup * kHalves);                               // state tiles of this CTA
    const int n_mtiles = (int)((n_frames + kTileM - 1) / kTileM);
    auto valid_of = [&](int h) { return min(kStatesPerTile, n_states - (sup * kHalves + h) * kStatesPerTile); };

    // ---- one-time setup: barriers, TMEM, resident B tile
    if (tid == 0) {
        for (int i = 0; i < kRawBufs; ++i) mbar_init(&sm.raw_full[i], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.a_full[i], IMG ? 1 : kTileM);
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
        const uint4* src = reinterpret_cast<const uint4*>(b_packed + (size_t)sup * kHalves * kBBytes);
        uint4* dst = reinterpret_cast<uint4*>(&sm.b[0][0]);
        for (int i = tid; i < H * (kBBytes / 16); i += kThreads) dst[i] = __ldg(src + i);
        if (tid < H * kStatesPerTile) sm.cst[tid] = cst_pad[sup * kHalves * kStatesPerTile + tid];
        // the zero chunk of both A stages is written once
        for (int i = tid; i < 2 * kTileM; i += kThreads)
            *reinterpret_cast<uint4*>(sm.a[i / kTileM] + (kAChunks - 1) * kALbo + (i % kTileM) * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    if (warp < kProducerThreads / 32) {
        // =========================== producers ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProducerRegs));
        const int n_it = (g < n_mtiles) ? (n_mtiles - g + G - 1) / G : 0;
        if constexpr (IMG) {
            // the image is tile-major and padded to whole tiles: one bulk copy per tile, stage it & 1
            if (tid == 0) {
                const uint8_t* img = reinterpret_cast<const uint8_t*>(feat);
                for (int it = 0; it < n_it; ++it) {
                    const int st = it & 1;
                    mbar_wait(&sm.a_empty[st], (((uint32_t)(it >> 1)) & 1) ^ 1);       // MMA finished reading this stage
                    bulk_load(sm.a[st], img + (size_t)(g + it * G) * kImgTileBytes, kImgTileBytes, &sm.a_full[st]);
                }
            }
        } else {
        constexpr int kTileElems = kTileM * kDim;                       // 4992 floats = 19968 B, a multiple of 16
        constexpr uint32_t kTileBytes = kTileElems * sizeof(float);
        // bulk copies need a 16-byte aligned source (use_bulk) and a whole tile
        auto tile_full = [&](int it) { return use_bulk && (int64_t)(g + it * G + 1) * kTileM <= n_frames; };
        // kProducerGroups groups of 128 threads (thread = feature row) take the tiles in turn; tile it is staged into
        // stage it & 1 of A.  The raw tiles go round four buffers (it & 3): four bulk copies are in flight.
        const int row_id = tid & (kTileM - 1);
        const int p = tid >> 7;
        auto issue = [&](int it) {                       // one thread: full tiles are contiguous and 16-byte aligned
            bulk_load(sm.raw[it & (kRawBufs - 1)], feat + (int64_t)(g + it * G) * kTileElems, kTileBytes, &sm.raw_full[it & (kRawBufs - 1)]);
        };
        if (row_id == 0)
            for (int it = p; it < kRawBufs && it < n_it; it += kProducerGroups)
                if (tile_full(it)) issue(it);
        for (int it = p; it < n_it; it += kProducerGroups) {
            const int st = it & 1;
            const uint32_t k = (uint32_t)(it >> 1);
            float* raw = sm.raw[it & (kRawBufs - 1)];
            if (tile_full(it)) {
                mbar_wait(&sm.raw_full[it & (kRawBufs - 1)], (uint32_t)(it / kRawBufs) & 1);         // TMA bytes have landed
            } else {
                // the batch's last, partial tile: plain loads, rows beyond the end read as zero
                const int64_t f0 = (int64_t)(g + it * G) * kTileM;
                const int total = (int)(n_frames - f0) * kDim;
                for (int e = row_id; e < kTileElems; e += kTileM) raw[e] = (e < total) ? __ldg(feat + f0 * kDim + e) : 0.f;
                asm volatile("bar.sync %0, 128;" ::"r"(1 + p) : "memory");
            }
            mbar_wait(&sm.a_empty[st], (k & 1) ^ 1);       // MMA finished reading this stage
            sm.inv2[it & 3][row_id] = stage_row(raw + row_id * kDim /* stride 39 words: conflict free */,
                                                sm.a[st] + row_id * 16);
            fence_proxy_async();
            mbar_arrive(&sm.a_full[st]);
            asm volatile("bar.sync %0, 128;" ::"r"(1 + p) : "memory");   // the group is done with this raw buffer: refill it
            if (row_id == 0 && it + kRawBufs < n_it && tile_full(it + kRawBufs)) issue(it + kRawBufs);
        }
        }
    } else if (warp == (kProducerThreads + kEpilogueThreads) / 32) {
        // =========================== MMA issuer ===========================
        // The whole warp walks the loop in step (uniform control flow, every descriptor built once, before the loop, from
        // values ptxas can prove uniform) and the MMAs / commits are predicated on one elected lane INSIDE the asm: the
        // operands then live in uniform registers.  Issued from a divergent `if (lane == 0)` the same code cost ~18
        // instructions per MMA (R2UR moves and an ELECT / broadcast loop around every UTCHMMA), and that single thread's
        // instruction stream, not the tensor pipe, paced the kernel.
        {
            const uint32_t leader = elect_one();
            // c_format F32 (bit 4), a/b format F16 (0), N >> 3 at bit 17, M >> 4 at bit 24
            uint32_t idesc[kNumMma];
            uint64_t a_desc[2][kNumMma];
            uint64_t b_desc[kHalves][kNumMma];
#pragma unroll
            for (int i = 0; i < kNumMma; ++i)
                idesc[i] = (1u << 4) | ((uint32_t)((pair_chunk(i, 4) * kBlockCols) >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
#pragma unroll
            for (int st = 0; st < 2; ++st) {
                const uint32_t a_base = smem_u32(sm.a[st]);
#pragma unroll
                for (int i = 0; i < kNumMma; ++i)
                    a_desc[st][i] = make_desc(a_base + pair_chunk(i, 0) * kALbo, (pair_chunk(i, 1) - pair_chunk(i, 0)) * kALbo);
            }
#pragma unroll
            for (int h = 0; h < kHalves; ++h) {
                const uint32_t b_base = smem_u32(sm.b[h]);
#pragma unroll
                for (int i = 0; i < kNumMma; ++i)
                    b_desc[h][i] = make_desc(b_base + pair_chunk(i, 2) * kBLbo, (pair_chunk(i, 3) - pair_chunk(i, 2)) * kBLbo);
            }
            // j counts (feature tile, half) pairs: accumulator j & 1, A stage = tile & 1
            int it = 0, j = 0;
            for (int m = g; m < n_mtiles; m += G, ++it) {
                const int st = it & 1;
                mbar_wait(&sm.a_full[st], (uint32_t)(it >> 1) & 1);
#pragma unroll
                for (int h = 0; h < kHalves; ++h) {
                    if (h >= H) break;
                    const int s = j & 1;
                    mbar_wait(&sm.tmem_empty[s], ((uint32_t)(j >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)(s * kBufStride);
#pragma unroll
                    for (int i = 0; i < kNumMma; ++i)
                        mma_f16_elected(leader, d, st ? a_desc[1][i] : a_desc[0][i], b_desc[h][i], idesc[i], i ? 1u : 0u);
                    if (h == H - 1) mma_commit_elected(leader, &sm.a_empty[st]);     // the stage is free once this tile's last product has run
                    mma_commit_elected(leader, &sm.tmem_full[s]);
                    ++j;
                }
            }
        }
    } else {
        // =========================== epilogue ===========================
        // One warp per TMEM lane quarter (thread = frame row).  The stores of a tile are deferred until the first
        // TMEM loads of the next tile have been issued.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpilogueRegs));
        // 8-byte stores of score pairs need an even row pitch and an 8-byte aligned matrix (and an even number of states
        // in the tile, else the odd last state would be lost): otherwise scalar stores
        const bool pair_ok = ((ld_out & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
        const int q = warp & 3;                             // TMEM lane quarter this warp may touch
        const int r = q * 32 + lane;
        float2* stage = reinterpret_cast<float2*>(sm.out_stage[q]);
        auto store_tile = [&](int m_prev, int h_prev) {
            const int st_base = (sup * kHalves + h_prev) * kStatesPerTile;
            const int valid = valid_of(h_prev);
            if (pair_ok && (valid & 1) == 0) {
                // a row's 6 scores are 24 contiguous bytes of the score matrix: three neighbouring lanes write one
                // row (8 bytes each) -- a third of the sector requests of lane-per-row scalar stores
                const int64_t f0 = (int64_t)m_prev * kTileM + q * 32;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int idx = i * 32 + lane, row = idx / 3, c = idx - row * 3;
                    const int st0 = st_base + 2 * c;
                    if (f0 + row < n_frames && st0 + 1 < n_states)
                        *reinterpret_cast<float2*>(out + (f0 + row) * ld_out + st0) = stage[idx];
                }
            } else {
                const int64_t f = (int64_t)m_prev * kTileM + r;
                if (f < n_frames) {
                    float* o = out + f * ld_out + st_base;
                    const float* mine = reinterpret_cast<const float*>(stage) + lane * kStatesPerTile;
#pragma unroll
                    for (int jj = 0; jj < kStatesPerTile; ++jj)
                        if (jj < valid) o[jj] = mine[jj];
                }
            }
            __syncwarp();                                   // the staging buffer may be rewritten
        };
        // j counts (feature tile, half) pairs like the MMA warp does: accumulator j & 1
        int it = 0, j = 0, m_prev = -1, h_prev = 0;
        for (int m = g; m < n_mtiles; m += G, ++it) {
#pragma unroll
            for (int h = 0; h < kHalves; ++h) {
                if (h >= H) break;
                const int s = j & 1;
                const int valid = valid_of(h);
                mbar_wait(&sm.tmem_full[s], (uint32_t)(j >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * kBufStride);
                // Two rounds of three states (120 accumulator columns each): what a round costs is the issue -> wait::ld
                // round trip (a couple of hundred cycles with four warps draining), not the bytes, so an accumulator is
                // drained in two waits instead of six and handed back to the MMA warp as soon as the second round sits
                // in registers, before any of its arithmetic.
                // Two rounds of 120 accumulator columns; column n belongs to state_of(n) (pieces of 8 columns, so a pair of
                // neighbouring columns never straddles two states).  The padding column of every state is an
                // exact zero: all 40 are summed.
                float v[kTileN / 2];
                auto load_half = [&](uint32_t t) {
                    tmem_ld64(t, v);
                    tmem_ld32(t + 64, v + 64);
                    tmem_ld16(t + 96, v + 96);
                    tmem_ld8(t + 112, v + 112);
                };
                const float* cst = sm.cst + h * kStatesPerTile;
                unsigned long long acc[kStatesPerTile];
#pragma unroll
                for (int c = 0; c < kStatesPerTile; ++c) acc[c] = 0ull;
                load_half(taddr);
                if (m_prev >= 0) store_tile(m_prev, h_prev);     // previous scores: shared memory -> global
                // exact: inv2 is a power of two (1 for ordinary rows)
                const float mhalf = -0.5f * (IMG ? __ldg(inv2_g + (size_t)m * kTileM + r) : sm.inv2[it & 3][r]);
                tmem_ld_wait();
#pragma unroll
                for (int n = 0; n < kTileN / 2; n += 2) acc[state_of(n)] = ffma2(pack_f2(v[n], v[n + 1]), acc[state_of(n)]);
                load_half(taddr + kTileN / 2);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&sm.tmem_empty[s]);
#pragma unroll
                for (int n = 0; n < kTileN / 2; n += 2)
                    acc[state_of(kTileN / 2 + n)] = ffma2(pack_f2(v[n], v[n + 1]), acc[state_of(kTileN / 2 + n)]);
                float score[kStatesPerTile];
#pragma unroll
                for (int c = 0; c < kStatesPerTile; ++c) {
                    const float2 f = unpack_f2(acc[c]);
                    score[c] = (c < valid) ? fmaf(mhalf, f.x + f.y, cst[c]) : 0.f;
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) stage[lane * 3 + c] = make_float2(score[2 * c], score[2 * c + 1]);
                __syncwarp();
                m_prev = m;
                h_prev = h;
                ++j;
            }
        }
        if (m_prev >= 0) store_tile(m_prev, h_prev);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kProducerThreads / 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

__global__ void __launch_bounds__(kThreads, 1)
emission_h16_kernel(const float* __restrict__ feat, int64_t n_frames, const uint8_t* __restrict__ b_packed,
                   const float* __restrict__ cst_pad, int n_states, float* __restrict__ out, int ld_out, int use_bulk,
                   int g_full, int g_last) {
    emission_h16_body<false>(feat, nullptr, n_frames, b_packed, cst_pad, n_states, out, ld_out, use_bulk, g_full, g_last, (int)blockIdx.x);
}

__global__ void __launch_bounds__(kThreads, 1)
emission_h16_img_kernel(const uint8_t* __restrict__ a_img, const float* __restrict__ inv2, int64_t n_frames, const uint8_t* __restrict__ b_packed,
                        const float* __restrict__ cst_pad, int n_states, float* __restrict__ out, int ld_out, int g_full, int g_last) {
    emission_h16_body<true>(reinterpret_cast<const float*>(a_img), inv2, n_frames, b_packed, cst_pad, n_states, out, ld_out, 1, g_full, g_last,
                            (int)blockIdx.x);
}

// Several models in ONE launch (batched training: one word model per segment, hidden_markov_model.py:294-318 runs them one
// after the other).  Segment i scores the frames [seg_begin[i], seg_end[i]) with the n_states[i] <= 12 states whose image
// starts at tile seg_tile[i] into the columns from seg_col[i]; it owns ctas_per_seg consecutive CTAs.  A segment whose
// active flag is not 1 is skipped by its CTAs (the device-side M-step freezes converged models, mstep.cu).
__global__ void __launch_bounds__(kThreads, 1)
emission_h16_multi_kernel(const float* __restrict__ feat, const uint8_t* __restrict__ b_packed, const float* __restrict__ cst_pad,
                          const int64_t* __restrict__ seg_begin, const int64_t* __restrict__ seg_end, const int32_t* __restrict__ seg_tile,
                          const int32_t* __restrict__ seg_states, const int32_t* __restrict__ seg_col, const int32_t* __restrict__ active,
                          float* __restrict__ out, int ld_out, int ctas_per_seg) {
    const int seg = blockIdx.x / ctas_per_seg, g = blockIdx.x % ctas_per_seg;
    if (active && active[seg] != 1) return;
    const int64_t begin = seg_begin[seg], n = seg_end[seg] - begin;
    const int n_mtiles = (int)((n + kTileM - 1) / kTileM);
    if (n <= 0 || g >= n_mtiles) return;
    const float* f = feat + begin * kDim;
    const int use_bulk = (reinterpret_cast<uintptr_t>(f) & 15) == 0 ? 1 : 0;
    emission_h16_body<false>(f, nullptr, n, b_packed + (size_t)seg_tile[seg] * kBBytes, cst_pad + seg_tile[seg] * kStatesPerTile, seg_states[seg],
                      out + begin * ld_out + seg_col[seg], ld_out, use_bulk, 1, min(ctas_per_seg, n_mtiles), g);
}

// The multi-model launch from PRE-SPLIT images (training: the features never change, so their A operand is built once by
// h16_image_kernel and every iteration only bulk-copies it).  Segment i owns the image tiles from seg_img_tile[i]
// (ceil(frames / 128) of them, rows beyond the segment zero) and the rows of inv2 from 128 * seg_img_tile[i].
__global__ void __launch_bounds__(kThreads, 1)
emission_h16_multi_img_kernel(const uint8_t* __restrict__ a_img, const float* __restrict__ inv2, const uint8_t* __restrict__ b_packed,
                              const float* __restrict__ cst_pad, const int64_t* __restrict__ seg_begin, const int64_t* __restrict__ seg_end,
                              const int32_t* __restrict__ seg_img_tile, const int32_t* __restrict__ seg_tile,
                              const int32_t* __restrict__ seg_states, const int32_t* __restrict__ seg_col, const int32_t* __restrict__ active,
                              float* __restrict__ out, int ld_out, int ctas_per_seg) {
    const int seg = blockIdx.x / ctas_per_seg, g = blockIdx.x % ctas_per_seg;
    if (active && active[seg] != 1) return;
    const int64_t begin = seg_begin[seg], n = seg_end[seg] - begin;
    const int n_mtiles = (int)((n + kTileM - 1) / kTileM);
    if (n <= 0 || g >= n_mtiles) return;
    const size_t t0 = (size_t)seg_img_tile[seg];
    emission_h16_body<true>(reinterpret_cast<const float*>(a_img + t0 * kImgTileBytes), inv2 + t0 * kTileM, n,
                            b_packed + (size_t)seg_tile[seg] * kBBytes, cst_pad + seg_tile[seg] * kStatesPerTile, seg_states[seg],
                            out + begin * ld_out + seg_col[seg], ld_out, 1, 1, min(ctas_per_seg, n_mtiles), g);
}

// float32 features -> pre-split A image, one thread per image row (segment-relative tiles; rows past a segment's end: zero)
__global__ void __launch_bounds__(128)
h16_image_kernel(const float* __restrict__ feat, const int64_t* __restrict__ seg_begin, const int64_t* __restrict__ seg_end,
                 const int32_t* __restrict__ seg_img_tile, int n_seg, uint8_t* __restrict__ a_img, float* __restrict__ inv2) {
    const int tile = blockIdx.x, row = threadIdx.x;
    int lo = 0, hi = n_seg;                              // segment of this tile: last one with seg_img_tile <= tile
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (seg_img_tile[mid] <= tile) lo = mid; else hi = mid; }
    const int64_t local = (int64_t)(tile - seg_img_tile[lo]) * kTileM + row;
    const int64_t f = seg_begin[lo] + local;
    uint8_t* a_row = a_img + (size_t)tile * kImgTileBytes + (size_t)row * 16;
    if (f < seg_end[lo]) {
        float v[kDim];
#pragma unroll
        for (int c = 0; c < kDim; ++c) v[c] = __ldg(feat + f * kDim + c);
        inv2[(size_t)tile * kTileM + row] = stage_row(v, a_row);
    } else {
#pragma unroll
        for (int c = 0; c < 2 * kStageChunksPerPart; ++c) *reinterpret_cast<uint4*>(a_row + c * kStageLbo) = make_uint4(0, 0, 0, 0);
        inv2[(size_t)tile * kTileM + row] = 1.0f;
    }
}

static_assert(sizeof(Smem) <= 227 * 1024, "shared memory of the emission kernel");

}  // namespace h16
}  // namespace loe

extern "C" int loe_emission_h16_tile_bytes(void) { return loe::h16::kBBytes; }

extern "C" int loe_emission_h16_multi_dev(const float* feat_dev, int dim, const void* b_packed_dev, const float* cst_pad_dev, int n_seg,
                                          const int64_t* seg_begin_dev, const int64_t* seg_end_dev, const int32_t* seg_tile_dev,
                                          const int32_t* seg_states_dev, const int32_t* seg_col_dev, const int32_t* active_dev,
                                          int max_states, float* out_dev, int ld_out, void* stream) {
    using namespace loe;
    using namespace loe::h16;
    if (n_seg <= 0) return LOE_OK;
    if (dim != kDim) { set_error("tensor-core emission path is built for dim == 39 (got %d)", dim); return LOE_ERR_UNSUPPORTED; }
    if (max_states > kHalves * kStatesPerTile) {
        set_error("a segment of the multi-model launch holds at most %d states (got %d)", kHalves * kStatesPerTile, max_states);
        return LOE_ERR_UNSUPPORTED;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    int dev = 0, sms = 0;
    LOE_CUDA(cudaGetDevice(&dev));
    LOE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static bool attr_done[64] = {false};
    if (dev < 64 && !attr_done[dev]) {
        LOE_CUDA(cudaFuncSetAttribute(emission_h16_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        attr_done[dev] = true;
    }
    // one CTA per SM (the kernel's shared memory allows no more), shared evenly by the segments (their frame counts are
    // similar in training); more segments than SMs simply queue
    int per_seg = sms / n_seg;
    if (per_seg < 1) per_seg = 1;
    emission_h16_multi_kernel<<<(unsigned)(n_seg * per_seg), kThreads, sizeof(Smem), s>>>(
        feat_dev, static_cast<const uint8_t*>(b_packed_dev), cst_pad_dev, seg_begin_dev, seg_end_dev, seg_tile_dev, seg_states_dev,
        seg_col_dev, active_dev, out_dev, ld_out, per_seg);
    LOE_LAUNCH_CHECK("emission_h16_multi_kernel");
    return LOE_OK;
}

namespace loe {
namespace h16 {
// divide the SMs over the column supertiles in proportion to their MMA cost (the last one may be narrower)
static int plan_grid(int n_states, int64_t n_frames, int* g_full_out, int* g_last_out, unsigned* grid_out) {
    static int sm_count[64] = {0};
    int dev = 0;
    LOE_CUDA(cudaGetDevice(&dev));
    if (dev >= 64) dev = 63;
    if (sm_count[dev] == 0) {
        LOE_CUDA(cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev));
        LOE_CUDA(cudaFuncSetAttribute(emission_h16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        LOE_CUDA(cudaFuncSetAttribute(emission_h16_img_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    }
    const int n_tiles = (n_states + kStatesPerTile - 1) / kStatesPerTile;
    const int n_super = (n_tiles + kHalves - 1) / kHalves;
    const int n_mtiles = (int)((n_frames + kTileM - 1) / kTileM);
    double cost_last = 0.0;
    for (int t = (n_super - 1) * kHalves; t < n_tiles; ++t) {
        const int valid = n_states - t * kStatesPerTile < kStatesPerTile ? n_states - t * kStatesPerTile : kStatesPerTile;
        cost_last += (double)(((valid * kColsPerState + 15) / 16) * 16) / (double)(kHalves * kTileN);
    }
    int g_full = 1, g_last = 1;
    const int sms = sm_count[dev];
    if (n_super == 1) {
        g_last = sms;
    } else if (sms >= n_super) {
        // minimise max(1 / g_full, cost_last / g_last) subject to (n_super - 1) * g_full + g_last <= sms
        double best = 1e30;
        for (int gf = 1; (n_super - 1) * gf < sms; ++gf) {
            const int gl = sms - (n_super - 1) * gf;
            const double t = (1.0 / gf > cost_last / gl) ? 1.0 / gf : cost_last / gl;
            if (t < best) { best = t; g_full = gf; g_last = gl; }
        }
    }
    if (g_full > n_mtiles) g_full = n_mtiles;
    if (g_last > n_mtiles) g_last = n_mtiles;
    *g_full_out = g_full; *g_last_out = g_last;
    *grid_out = (unsigned)((n_super - 1) * g_full + g_last);
    return LOE_OK;
}
}  // namespace h16
}  // namespace loe

extern "C" int loe_emission_h16_dev(const float* feat_dev, int64_t n_frames, int dim, const void* b_packed_dev,
                                    const float* cst_pad_dev, int n_states, float* out_dev, int ld_out, void* stream) {
    using namespace loe;
    using namespace loe::h16;
    if (n_frames <= 0 || n_states <= 0) return LOE_OK;
    if (dim != kDim) { set_error("tensor-core emission path is built for dim == 39 (got %d)", dim); return LOE_ERR_UNSUPPORTED; }
    if (ld_out < n_states) { set_error("ld_out (%d) < n_states (%d)", ld_out, n_states); return LOE_ERR_VALUE; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    int g_full = 1, g_last = 1; unsigned grid = 1;
    const int st = plan_grid(n_states, n_frames, &g_full, &g_last, &grid);
    if (st != LOE_OK) return st;
    const int use_bulk = (reinterpret_cast<uintptr_t>(feat_dev) & 15) == 0 ? 1 : 0;
    emission_h16_kernel<<<grid, kThreads, sizeof(Smem), s>>>(feat_dev, n_frames, static_cast<const uint8_t*>(b_packed_dev), cst_pad_dev,
                                                             n_states, out_dev, ld_out, use_bulk, g_full, g_last);
    LOE_LAUNCH_CHECK("emission_h16_kernel");
    return LOE_OK;
}

extern "C" int loe_h16_image_dev(const float* feat_dev, int dim, int n_seg, const int64_t* seg_begin_dev, const int64_t* seg_end_dev,
                                 const int32_t* seg_img_tile_dev, int n_img_tiles, void* a_img_dev, float* inv2_dev, void* stream) {
    using namespace loe;
    using namespace loe::h16;
    if (n_seg <= 0 || n_img_tiles <= 0) return LOE_OK;
    if (dim != kDim) { set_error("tensor-core emission path is built for dim == 39 (got %d)", dim); return LOE_ERR_UNSUPPORTED; }
    if ((reinterpret_cast<uintptr_t>(a_img_dev) & 15) != 0) { set_error("a_img_dev must be 16-byte aligned"); return LOE_ERR_VALUE; }
    h16_image_kernel<<<(unsigned)n_img_tiles, kTileM, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        feat_dev, seg_begin_dev, seg_end_dev, seg_img_tile_dev, n_seg, static_cast<uint8_t*>(a_img_dev), inv2_dev);
    LOE_LAUNCH_CHECK("h16_image_kernel");
    return LOE_OK;
}

extern "C" int loe_emission_h16_multi_img_dev(const void* a_img_dev, const float* inv2_dev, const void* b_packed_dev, const float* cst_pad_dev,
                                              int n_seg, const int64_t* seg_begin_dev, const int64_t* seg_end_dev,
                                              const int32_t* seg_img_tile_dev, const int32_t* seg_tile_dev, const int32_t* seg_states_dev,
                                              const int32_t* seg_col_dev, const int32_t* active_dev, int max_states, float* out_dev,
                                              int ld_out, void* stream) {
    using namespace loe;
    using namespace loe::h16;
    if (n_seg <= 0) return LOE_OK;
    if (max_states > kHalves * kStatesPerTile) {
        set_error("a segment of the multi-model launch holds at most %d states (got %d)", kHalves * kStatesPerTile, max_states);
        return LOE_ERR_UNSUPPORTED;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    int dev = 0, sms = 0;
    LOE_CUDA(cudaGetDevice(&dev));
    LOE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static bool attr_done[64] = {false};
    if (dev < 64 && !attr_done[dev]) {
        LOE_CUDA(cudaFuncSetAttribute(emission_h16_multi_img_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        attr_done[dev] = true;
    }
    int per_seg = sms / n_seg;
    if (per_seg < 1) per_seg = 1;
    emission_h16_multi_img_kernel<<<(unsigned)(n_seg * per_seg), kThreads, sizeof(Smem), s>>>(
        static_cast<const uint8_t*>(a_img_dev), inv2_dev, static_cast<const uint8_t*>(b_packed_dev), cst_pad_dev, seg_begin_dev, seg_end_dev,
        seg_img_tile_dev, seg_tile_dev, seg_states_dev, seg_col_dev, active_dev, out_dev, ld_out, per_seg);
    LOE_LAUNCH_CHECK("emission_h16_multi_img_kernel");
    return LOE_OK;
}

extern "C" int64_t loe_emission_h16_img_bytes(int64_t n_frames) {
    return ((n_frames + loe::h16::kTileM - 1) / loe::h16::kTileM) * (int64_t)loe::h16::kImgTileBytes;
}

extern "C" int loe_emission_h16_img_dev(const void* a_img_dev, const float* inv2_dev, int64_t n_frames, const void* b_packed_dev,
                                        const float* cst_pad_dev, int n_states, float* out_dev, int ld_out, void* stream) {
    using namespace loe;
    using namespace loe::h16;
    if (n_frames <= 0 || n_states <= 0) return LOE_OK;
    if (ld_out < n_states) { set_error("ld_out (%d) < n_states (%d)", ld_out, n_states); return LOE_ERR_VALUE; }
    if ((reinterpret_cast<uintptr_t>(a_img_dev) & 15) != 0) { set_error("a_img_dev must be 16-byte aligned (bulk copies)"); return LOE_ERR_VALUE; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    int g_full = 1, g_last = 1; unsigned grid = 1;
    const int st = plan_grid(n_states, n_frames, &g_full, &g_last, &grid);
    if (st != LOE_OK) return st;
    emission_h16_img_kernel<<<grid, kThreads, sizeof(Smem), s>>>(static_cast<const uint8_t*>(a_img_dev), inv2_dev, n_frames,
                                                                 static_cast<const uint8_t*>(b_packed_dev), cst_pad_dev, n_states, out_dev,
                                                                 ld_out, g_full, g_last);
    LOE_LAUNCH_CHECK("emission_h16_img_kernel");
    return LOE_OK;
}
