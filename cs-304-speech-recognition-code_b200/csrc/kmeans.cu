// Segmental K-means sufficient statistics for sm_100a.
//
// Replaces the accumulation half of the reference's M-step:
//   Signal.order_by_state                      signal.py:23-47   (contiguous run per state, increasing order)
//   SortedSignals.transition_probabilities     signal.py:81-91   (consecutive-pair counts)
//   _update_middleware_parameters              hidden_markov_model.py:320-350 (per-state mean / np.cov inputs)
//   HiddenMarkovModelMultiWord._remux_path_and_signal  hidden_markov_model.py:602-636 (cut the chain alignment
//       per word label, re-base to the word's first state, final piece never flushed)
//
// align_kernel:  one thread per utterance walks the alignment once, writes for every frame the
//                bucket (global state id of the word model it is credited to, 0xFFFF = none) and
//                adds integer transition counts with atomics (integer => order independent).
// accum_kernel:  grid = (bucket, chunk of frames).  A CTA scans the bucket ids of its chunk,
//                compacts the matching frames in order and accumulates the outer products of the
//                shifted, augmented frames [x - shift_g, 1] in 4 x 4 register tiles (float64); the
//                packed statistics vector is [N | sum x | upper triangle of sum x x^T].
//                No floating-point atomics: partials per chunk, then
// reduce_kernel: sums the chunk partials in a fixed order => bitwise reproducible statistics.
// Algorithmic HBM bytes: 4*D + 2 per frame (features once + bucket id; re-scans of the id array hit L2).
#include "common.cuh"
#include <stdlib.h>

namespace loe {

constexpr int kAccThreads = 256;
constexpr int kMaxDimK = 40;               // D + 1 (constant-one column) must fit

// ------------------------------------------------------------------------------------------
__global__ void align_kernel(const int8_t* __restrict__ path, const int64_t* __restrict__ frm_off, int n_utt,
                             const int32_t* __restrict__ tr_off, const int32_t* __restrict__ col,
                             const int32_t* __restrict__ word, const int32_t* __restrict__ word_lo,
                             const int32_t* __restrict__ utt_tr, int remux, int n_glob,
                             uint16_t* __restrict__ bucket, int32_t* __restrict__ counts) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_utt) return;
    const int64_t f0 = frm_off[u];
    const int T = (int)(frm_off[u + 1] - f0);
    const int tr = utt_tr ? utt_tr[u] : 0;
    const int p0 = tr_off[tr];
    const int P = tr_off[tr + 1] - p0;
    const int8_t* __restrict__ pth = path + f0;
    uint16_t* __restrict__ bk = bucket + f0;
    for (int t = 0; t < T; ++t) bk[t] = 0xFFFF;
    if (T <= 0 || pth[0] < 0 || pth[0] >= P) return;     // T == 1 gives path [-1]: nothing is credited

    int seg_start = 0;
    while (seg_start < T) {
        // segment = maximal run of frames whose position carries the same word label
        const int pfirst = pth[seg_start];
        if (pfirst < 0 || pfirst >= P) break;
        int seg_end = T;
        if (remux) {
            const int lab = word[p0 + pfirst];
            seg_end = seg_start + 1;
            while (seg_end < T) {
                const int p = pth[seg_end];
                if (p < 0 || p >= P || word[p0 + p] != lab) break;
                ++seg_end;
            }
            if (seg_end == T) break;                      // final piece is never flushed (:614-636)
        }
        const int lo = remux ? word_lo[p0 + pfirst] : 0;
        int n_states = P;
        if (remux) { n_states = 1; while (lo + n_states < P && word_lo[p0 + lo + n_states] == lo) ++n_states; }
        const int gbase = col[p0 + lo];
        // order_by_state: accept frames while the local state sequence is non-decreasing in [0, n)
        int last = 0; bool ok = true;
        int prev_local = 0;
        for (int t = seg_start; t < seg_end; ++t) {
            const int local = (int)pth[t] - lo;
            const bool in_range = local >= 0 && local < n_states;
            if (ok && in_range && local >= last) { bk[t] = (uint16_t)(gbase + local); last = local; }
            else ok = false;
            if (t > seg_start && in_range && prev_local >= 0 && prev_local < n_states)
                atomicAdd(counts + (size_t)(gbase + prev_local) * n_glob + (gbase + local), 1);
            prev_local = local;
        }
        seg_start = seg_end;
    }
}

// Isolated training (remux == 0): one WARP per utterance, lanes over the frames.  The serial walk above reduces to two
// prefix conditions -- a frame is credited while every frame so far is in range and the state sequence has not decreased
// -- and the transition counts to one integer atomic per DISTINCT (from, to) pair of 32 consecutive frames (lanes with the
// same pair elect a leader), instead of one per frame on a few dozen hot counters.
__global__ void __launch_bounds__(128)
align_warp_kernel(const int8_t* __restrict__ path, const int64_t* __restrict__ frm_off, int n_utt,
                  const int32_t* __restrict__ tr_off, const int32_t* __restrict__ col, const int32_t* __restrict__ utt_tr, int n_glob,
                  uint16_t* __restrict__ bucket, int32_t* __restrict__ counts) {
    constexpr unsigned FULL = 0xffffffffu;
    const int u = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (u >= n_utt) return;
    const int64_t f0 = frm_off[u];
    const int T = (int)(frm_off[u + 1] - f0);
    const int tr = utt_tr ? utt_tr[u] : 0;
    const int p0 = tr_off[tr];
    const int P = tr_off[tr + 1] - p0;
    const int gbase = col[p0];
    if (T <= 0) return;
    const int first = (int)path[f0];
    if (first < 0 || first >= P) {  // T == 1 gives path [-1]: nothing is credited, nothing counted (like the serial walk)
        for (int t = lane; t < T; t += 32) bucket[f0 + t] = (uint16_t)0xFFFF;
        return;
    }
    bool ok = true;                 // every frame before this round was credited
    int carry = 0;                  // local state of the last frame of the previous round ("last" starts at 0)
    bool carry_in = true;           // ... and whether it was in range (for the transition count of the round's first frame)
    for (int tb = 0; tb < T; tb += 32) {
        const int t = tb + lane;
        const int local = t < T ? (int)path[f0 + t] : 0;
        const bool in_range = t < T && local >= 0 && local < P;
        int prev = __shfl_up_sync(FULL, local, 1);
        bool prev_in = __shfl_up_sync(FULL, (int)in_range, 1) != 0;
        if (lane == 0) { prev = carry; prev_in = carry_in; }
        const bool good = in_range && local >= ((t == 0) ? 0 : prev);
        const unsigned bad = __ballot_sync(FULL, t < T && !good);
        const bool credited = ok && t < T && (bad == 0 || lane < __ffs(bad) - 1);
        if (t < T) bucket[f0 + t] = credited ? (uint16_t)(gbase + local) : (uint16_t)0xFFFF;
        ok = ok && bad == 0;
        // transition (prev -> local) for t >= 1 with both ends in range
        const bool count = t < T && t > 0 && in_range && prev_in;
        const int key = count ? prev * P + local : -1;
        const unsigned same = __match_any_sync(FULL, key);
        if (count && lane == __ffs(same) - 1) atomicAdd(counts + (size_t)(gbase + prev) * n_glob + (gbase + local), __popc(same));
        carry = __shfl_sync(FULL, local, 31);
        carry_in = __shfl_sync(FULL, (int)in_range, 31) != 0;
    }
}

// ------------------------------------------------------------------------------------------
// Outer-product accumulation, register tiled.  The augmented frame y = [x - shift, 1] has K = D + 1 <= 40
// entries; sum y y^T is cut into 4 x 4 tiles of a 10 x 10 tile grid and only the 55 tiles on or above the
// diagonal are computed.  The CTA's 256 threads form 4 groups of 64; group q takes frames q, q+4, ... of a
// batch and thread t < 55 of a group owns one tile (16 float64 accumulators): 4 + 4 shared-memory loads
// feed 16 FMAs.  Groups are summed in a fixed order at the end, chunks by reduce_kernel: reproducible.
constexpr int kAccGroups = 4;
constexpr int kAccGroupThreads = kAccThreads / kAccGroups;      // 64
constexpr int kTilesPerDim = kMaxDimK / 4;                       // 10
constexpr int kTiles = kTilesPerDim * (kTilesPerDim + 1) / 2;    // 55
constexpr int kAccFrames = 32;                                   // frames staged per batch

constexpr int kIdsPerThread = 8;
constexpr int kScanSpan = kAccThreads * kIdsPerThread;           // 2048 frames per scan round

// min / max bucket id present in every chunk: lets the (bucket, chunk) CTAs that have nothing to do exit
// at once (frames are grouped by word, so a chunk usually holds the states of one or two words)
__global__ void chunk_range_kernel(const uint16_t* __restrict__ bucket, int64_t total_frames, int64_t chunk,
                                   int* __restrict__ range) {
    __shared__ int s_lo[32], s_hi[32];
    const int c = blockIdx.x;
    const int64_t f_begin = (int64_t)c * chunk, f_end = min(total_frames, f_begin + chunk);
    int lo = 0x7fffffff, hi = -1;
    for (int64_t f = f_begin + threadIdx.x; f < f_end; f += blockDim.x) {
        const int b = bucket[f];
        if (b != 0xFFFF) { lo = min(lo, b); hi = max(hi, b); }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = min(lo, s_lo[w]); hi = max(hi, s_hi[w]); }
        range[2 * c] = lo; range[2 * c + 1] = hi;
    }
}

__global__ void __launch_bounds__(kAccThreads)
accum_kernel(const float* __restrict__ feat, const uint16_t* __restrict__ bucket, int64_t total_frames, int dim,
             int64_t chunk, const float* __restrict__ shift, const int* __restrict__ range, double* __restrict__ part) {
    const int g = blockIdx.x;
    const int c = blockIdx.y;
    if (g < range[2 * c] || g > range[2 * c + 1]) return;          // partials were zeroed by the launcher
    const int n_chunks = gridDim.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = tid / kAccGroupThreads, tg = tid % kAccGroupThreads;
    const int stride = 1 + dim + dim * (dim + 1) / 2;

    __shared__ __align__(16) double s_x[kAccFrames][kMaxDimK];
    __shared__ float s_shift[kMaxDimK];
    __shared__ int s_list[kScanSpan];                             // matching frames of the round, relative to `base`
    __shared__ int s_wcount[kAccThreads / 32];
    __shared__ double s_sum[kMaxDimK][kMaxDimK + 1];

    // tile of this thread: t -> (ti, tj), ti <= tj, row-major over the upper triangle
    int ti = 0, tj = 0;
    const bool has_tile = tg < kTiles;
    if (has_tile) {
        int r = tg;
        while (r >= kTilesPerDim - ti) { r -= kTilesPerDim - ti; ++ti; }
        tj = ti + r;
    }
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    if (tid < kMaxDimK) s_shift[tid] = (tid < dim) ? shift[(size_t)g * dim + tid] : 0.f;
    __syncthreads();

    const int64_t f_begin = (int64_t)c * chunk;
    const int64_t f_end = min(total_frames, f_begin + chunk);
    for (int64_t base = f_begin; base < f_end; base += kScanSpan) {
        // ---- scan 8 consecutive ids per thread, compact the matches in frame order
        const int64_t f0 = base + (int64_t)tid * kIdsPerThread;
        unsigned flags = 0;
        if (f0 + kIdsPerThread <= f_end && ((reinterpret_cast<uintptr_t>(bucket + f0) & 15) == 0)) {
            const uint4 v = *reinterpret_cast<const uint4*>(bucket + f0);
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if ((w[q] & 0xFFFFu) == (unsigned)g) flags |= 1u << (2 * q);
                if ((w[q] >> 16) == (unsigned)g) flags |= 1u << (2 * q + 1);
            }
        } else {
#pragma unroll
            for (int q = 0; q < kIdsPerThread; ++q)
                if (f0 + q < f_end && bucket[f0 + q] == (uint16_t)g) flags |= 1u << q;
        }
        const int cnt = __popc(flags);
        int incl = cnt;                                           // inclusive warp prefix sum
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_wcount[warp] = incl;
        __syncthreads();
        int off = incl - cnt, n_match = 0;
#pragma unroll
        for (int w = 0; w < kAccThreads / 32; ++w) { if (w < warp) off += s_wcount[w]; n_match += s_wcount[w]; }
#pragma unroll
        for (int q = 0; q < kIdsPerThread; ++q)
            if (flags & (1u << q)) s_list[off++] = tid * kIdsPerThread + q;
        __syncthreads();
        for (int b0 = 0; b0 < n_match; b0 += kAccFrames) {
            const int nb = min(kAccFrames, n_match - b0);
            // stage y = [x - shift, 1, 0...] of nb frames (float64)
            for (int i = tid; i < nb * kMaxDimK; i += kAccThreads) {
                const int r = i / kMaxDimK, k = i - r * kMaxDimK;
                double v = 0.0;
                if (k < dim) v = (double)feat[(base + s_list[b0 + r]) * dim + k] - (double)s_shift[k];
                else if (k == dim) v = 1.0;
                s_x[r][k] = v;
            }
            __syncthreads();
            if (has_tile) {
                for (int r = grp; r < nb; r += kAccGroups) {
                    const double2 i01 = *reinterpret_cast<const double2*>(&s_x[r][4 * ti]);
                    const double2 i23 = *reinterpret_cast<const double2*>(&s_x[r][4 * ti + 2]);
                    const double2 j01 = *reinterpret_cast<const double2*>(&s_x[r][4 * tj]);
                    const double2 j23 = *reinterpret_cast<const double2*>(&s_x[r][4 * tj + 2]);
                    const double xi[4] = {i01.x, i01.y, i23.x, i23.y};
                    const double xj[4] = {j01.x, j01.y, j23.x, j23.y};
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[a][b] = fma(xi[a], xj[b], acc[a][b]);
                }
            }
            __syncthreads();
        }
    }
    // groups -> one K x K matrix in shared memory (fixed order), then the packed vector
    for (int q = 0; q < kAccGroups; ++q) {
        if (grp == q && has_tile) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    double* dstp = &s_sum[4 * ti + a][4 * tj + b];
                    *dstp = (q == 0) ? acc[a][b] : (*dstp + acc[a][b]);
                }
        }
        __syncthreads();
    }
    double* dst = part + ((size_t)g * n_chunks + c) * stride;
    for (int e = tid; e < stride; e += kAccThreads) {
        int i = dim, j = dim;                              // e = 0: N = sum 1*1
        if (e >= 1 && e <= dim) { i = e - 1; j = dim; }    // sum (x_i - shift_i) * 1
        else if (e > dim) {
            int r = e - 1 - dim, row = 0;
            while (r >= dim - row) { r -= dim - row; ++row; }
            i = row; j = row + r;
        }
        dst[e] = s_sum[i][j];
    }
}

__global__ void reduce_kernel(const double* __restrict__ part, int n_chunks, int stride, double* __restrict__ stats) {
    const int g = blockIdx.x;
    for (int e = threadIdx.x; e < stride; e += blockDim.x) {
        double a = 0.0;
        const double* p = part + (size_t)g * n_chunks * stride + e;
        for (int c = 0; c < n_chunks; ++c) a += p[(size_t)c * stride];
        stats[(size_t)g * stride + e] = a;
    }
}

// ------------------------------------------------------------------------------------------
// Sorted path (dim == 39, n_glob <= 1024): the default.
//   1. counting sort of the frames by bucket, stable in frame order, per chunk of 2048 frames:
//        bucket_hist_kernel -> bucket_scan_kernel (exclusive offsets per (bucket, chunk)) -> bucket_plan_kernel (bases, work list)
//        -> bucket_scatter_kernel (frame index lists, one contiguous list per bucket)
//   2. accum2_kernel: one CTA per work item = up to 2048 consecutive list entries of ONE bucket.  The augmented frames
//      y = [x - shift, 1] are gathered through the list, 4 frames per warp and step, kA2Depth steps ahead of their use
//      (and the list entries one step earlier still), and sum y y^T = Y^T Y is a contraction over the frames on the FP64
//      TENSOR cores: mma.sync m8n8k4 (DMMA), 15 tiles of 8 x 8 on or above the diagonal of the 5 x 5 tile grid, 4 frames
//      per instruction, all 15 accumulator tiles in the registers of every warp; every lane gathers exactly the five
//      values that are its elements of the A and B fragments, so 5 global loads + 5 conversions feed 15 DMMAs (3 840 FMAs)
//      with no shared memory and no barrier in the loop.  The FP64 pipe bounds this kernel (960 FMAs per frame at 64 per
//      clock and SM: scratch/dmma_peak.cu measures 37 TFLOP/s); with scalar DFMAs in 8 x 8 register tiles (the first
//      version of this path) instruction issue and latency did: 15 % pipe utilisation (ncu).
//      float64 is kept because the models must come out identical whatever the number of ranks the frames are sharded
//      over -- split-precision products on the tcgen05 pipe give 2^-21 per product, not enough for that.
//   3. reduce2_kernel: partials of a bucket summed in work-list order.
// Everything is in a fixed order: statistics are bitwise reproducible, and independent of how other buckets' frames are
// interleaved with a bucket's own.
// ------------------------------------------------------------------------------------------
constexpr int kSortChunk = 2048;                 // frames per histogram / scatter CTA
constexpr int kSortThreads = 256;
constexpr int kSortMaxGlob = 1024;
constexpr int kSplit = 2048;                     // list entries per work item
#ifndef LOE_A2_DEPTH
#define LOE_A2_DEPTH 4
#endif
constexpr int kA2WarpDoubles = 15 * 2 * 32;       // accumulator values of one warp
constexpr int kA2Smem = 8 * kA2WarpDoubles * (int)sizeof(double);
constexpr int kA2Depth = LOE_A2_DEPTH;           // 4-frame steps each warp has in flight (gathered into registers ahead of use)

__global__ void __launch_bounds__(kSortThreads)
bucket_hist_kernel(const uint16_t* __restrict__ bucket, int64_t total_frames, int n_glob, int* __restrict__ chunk_hist) {
    extern __shared__ int s_hist[];
    const int c = blockIdx.x;
    for (int i = threadIdx.x; i < n_glob; i += kSortThreads) s_hist[i] = 0;
    __syncthreads();
    const int64_t f_begin = (int64_t)c * kSortChunk, f_end = min(total_frames, f_begin + kSortChunk);
    for (int64_t f = f_begin + threadIdx.x; f < f_end; f += kSortThreads) {
        const int b = bucket[f];
        if (b < n_glob) atomicAdd(&s_hist[b], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_glob; i += kSortThreads) chunk_hist[(size_t)i * gridDim.x + c] = s_hist[i];       // [bucket][chunk]
}

// warp per bucket: exclusive offsets over the chunks (in place, coalesced: the histogram is stored [bucket][chunk]) and the
// bucket's total
__global__ void __launch_bounds__(1024)
bucket_scan_kernel(int* __restrict__ chunk_hist, int n_chunks, int n_glob, int* __restrict__ total) {
    const int g = blockIdx.x * 32 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (g >= n_glob) return;
    int* h = chunk_hist + (size_t)g * n_chunks;
    int carry = 0;
    for (int c0 = 0; c0 < n_chunks; c0 += 32) {
        const int c = c0 + lane;
        const int v = (c < n_chunks) ? h[c] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (c < n_chunks) h[c] = carry + incl - v;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) total[g] = carry;
}

// one CTA: bucket bases (exclusive scan of the totals) and the work list -- up to kSplit consecutive list entries of one
// bucket per item, in bucket order
__global__ void __launch_bounds__(1024)
bucket_plan_kernel(const int* __restrict__ total, int n_glob, int* __restrict__ base, int* __restrict__ work,
                   int* __restrict__ work_range, int n_work_max) {
    __shared__ int s_base[kSortMaxGlob + 1], s_w0[kSortMaxGlob + 1];
    if (threadIdx.x == 0) {
        int b = 0, nw = 0;
        for (int q = 0; q < n_glob; ++q) {
            s_base[q] = b; s_w0[q] = nw;
            b += total[q];
            nw += (total[q] + kSplit - 1) / kSplit;
        }
        s_base[n_glob] = b; s_w0[n_glob] = min(nw, n_work_max);
    }
    __syncthreads();
    for (int q = threadIdx.x; q <= n_glob; q += blockDim.x) work_range[q] = min(s_w0[q], n_work_max);
    for (int q = threadIdx.x; q < n_glob; q += blockDim.x) {
        base[q] = s_base[q];
        const int n = total[q];
        for (int o = 0, w = s_w0[q]; o < n && w < n_work_max; o += kSplit, ++w) {
            work[3 * w] = q; work[3 * w + 1] = s_base[q] + o; work[3 * w + 2] = s_base[q] + min(n, o + kSplit);
        }
    }
}

__global__ void __launch_bounds__(kSortThreads)
bucket_scatter_kernel(const uint16_t* __restrict__ bucket, int64_t total_frames, int n_glob, const int* __restrict__ chunk_off,
                      const int* __restrict__ base, int* __restrict__ idx) {
    extern __shared__ int s_sc[];
    int* run_off = s_sc;                          // [n_glob] next free list slot of every bucket
    int* wc = s_sc + n_glob;                      // [8][n_glob] matches per warp in the current round
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n_glob; i += kSortThreads) run_off[i] = base[i] + chunk_off[(size_t)i * gridDim.x + c];
    for (int i = tid; i < 8 * n_glob; i += kSortThreads) wc[i] = 0;
    __syncthreads();
    const int64_t f_begin = (int64_t)c * kSortChunk;
    for (int r = 0; r < kSortChunk / kSortThreads; ++r) {
        const int64_t f = f_begin + r * kSortThreads + tid;
        const int b = (f < total_frames) ? (int)bucket[f] : 0xFFFF;
        const bool valid = b < n_glob;
        const unsigned mask = __match_any_sync(0xffffffffu, b);
        const int rank = __popc(mask & ((1u << lane) - 1u));
        const bool leader = valid && rank == 0;
        if (leader) wc[warp * n_glob + b] = __popc(mask);
        __syncthreads();
        if (valid) {
            int pos = run_off[b] + rank;
            for (int w = 0; w < warp; ++w) pos += wc[w * n_glob + b];
            idx[pos] = (int)(f - 0);              // frame index (total_frames < 2^31 is checked by the launcher)
        }
        __syncthreads();
        if (leader) { atomicAdd(&run_off[b], __popc(mask)); wc[warp * n_glob + b] = 0; }
        __syncthreads();
    }
}

// D (8 x 8, fp64) += A (8 x 4) B (4 x 8) on the FP64 tensor cores.  Lane l holds A[l >> 2][l & 3], B[l & 3][l >> 2] and
// D[l >> 2][2 (l & 3) + {0, 1}].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 2)
accum2_kernel(const float* __restrict__ feat, const int* __restrict__ idx, const int* __restrict__ work, const int* __restrict__ work_range,
              int n_glob, const float* __restrict__ shift, double* __restrict__ part) {
    constexpr int D = 39;
    const int item = blockIdx.x;
    if (item >= work_range[n_glob]) return;
    const int g = work[3 * item], begin = work[3 * item + 1], end = work[3 * item + 2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Every warp works on its own: steps of 4 frames (one DMMA K), step s of the item belongs to warp s % 8.  With Y the
    // 4 x 40 block of a step (rows y = [x - shift, 1]), tile (bi, bj) of Y^T Y takes A = (Y^T)[8 bi .., :] and B = Y[:, 8 bj ..],
    // and lane l's element of BOTH fragments is Y[l & 3][8 b + (l >> 2)]: a lane needs five values of ONE frame of the step
    // (columns (l >> 2) + 8 b), and it is the lane that loads them -- the fragments go from the gather straight into the
    // DMMAs without passing through shared memory.  A step's loads are issued kA2Depth steps of the warp ahead of their use,
    // and the frame-list entry they go through one step earlier still (the address of a gather never waits for the list).
    extern __shared__ double s_w[];                            // [8 warps][15 tiles x 2 halves][32 lanes], used after the loop
    double c[15][2];
#pragma unroll
    for (int t = 0; t < 15; ++t) { c[t][0] = 0.0; c[t][1] = 0.0; }
    const int row = lane & 3, col0 = lane >> 2;
    const bool one = col0 == 7;                                // block 4 of this lane is column 39: the constant 1
    double sh[5];
#pragma unroll
    for (int b = 0; b < 5; ++b) sh[b] = (col0 + 8 * b < D) ? (double)__ldg(shift + (size_t)g * D + col0 + 8 * b) : 0.0;
    const int n_steps = (end - begin + 3) >> 2;
    float pre[kA2Depth][5];
    int nidx[kA2Depth];
    // frame of this lane's row in step st (-1: past the item's end)
    auto list_entry = [&](int st) -> int {
        const int f = begin + 4 * st + row;
        return (st < n_steps && f < end) ? __ldg(idx + f) : -1;
    };
    auto gather = [&](int fi, float* dst) {
        const float* p = feat + (size_t)max(fi, 0) * D + col0;
        const bool on = fi >= 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) dst[b] = on ? __ldg(p + 8 * b) : 0.f;
        dst[4] = (on && !one) ? __ldg(p + 32) : 1.f;          // column 39 carries 1 (its shift entry is 0)
    };
#pragma unroll
    for (int d = 0; d < kA2Depth; ++d) nidx[d] = list_entry(warp + 8 * d);
#pragma unroll
    for (int d = 0; d < kA2Depth; ++d) { gather(nidx[d], pre[d]); nidx[d] = list_entry(warp + 8 * (d + kA2Depth)); }
    for (int st = warp; st < n_steps; st += 8 * kA2Depth) {
#pragma unroll
        for (int d = 0; d < kA2Depth; ++d) {
            const int cur = st + 8 * d;
            if (cur >= n_steps) break;                                     // warp-uniform
            double f[5];
#pragma unroll
            for (int b = 0; b < 5; ++b) f[b] = (double)pre[d][b] - sh[b];
            if (begin + 4 * cur + 3 >= end) {                              // warp-uniform: the item's last, partial step --
                const bool live = begin + 4 * cur + row < end;            // rows beyond the item are ZERO, also their constant 1
#pragma unroll
                for (int b = 0; b < 5; ++b) f[b] = live ? f[b] : 0.0;
            }
            gather(nidx[d], pre[d]);                                       // step cur + 8 kA2Depth
            nidx[d] = list_entry(cur + 16 * kA2Depth);
            int t = 0;
#pragma unroll
            for (int bi = 0; bi < 5; ++bi)
#pragma unroll
                for (int bj = bi; bj < 5; ++bj, ++t) dmma884(c[t][0], c[t][1], f[bi], f[bj]);
        }
    }
    // the eight warps' accumulators meet in shared memory ([warp][tile, half][lane]: conflict-free, one barrier) and every
    // output element is summed over the warps in warp order by the thread that stores it
    {
        double* mine = s_w + warp * kA2WarpDoubles;
        int t = 0;
#pragma unroll
        for (int bi = 0; bi < 5; ++bi)
#pragma unroll
            for (int bj = bi; bj < 5; ++bj, ++t) {
                mine[(2 * t) * 32 + lane] = c[t][0];
                mine[(2 * t + 1) * 32 + lane] = c[t][1];
            }
    }
    __syncthreads();
    constexpr int stride = 1 + D + D * (D + 1) / 2;
    double* dst = part + (size_t)item * stride;
    for (int e = tid; e < stride; e += 256) {
        int i = D, j = D;                                  // e = 0: N = sum 1 * 1
        if (e >= 1 && e <= D) { i = e - 1; j = D; }        // sum (x_i - shift_i) * 1
        else if (e > D) {
            int r = e - 1 - D, row = 0;
            while (r >= D - row) { r -= D - row; ++row; }
            i = row; j = row + r;
        }
        // element (i, j), i <= j, sits in tile (i / 8, j / 8) of the triangular tile order, in lane 4 (i % 8) + (j % 8) / 2, half j % 2
        const int bi = i >> 3, bj = j >> 3;
        const int t = bi * 5 - (bi * (bi - 1)) / 2 + (bj - bi);
        const int q = (2 * t + (j & 1)) * 32 + ((i & 7) << 2) + ((j & 7) >> 1);
        double a = s_w[q];
#pragma unroll
        for (int w = 1; w < 8; ++w) a += s_w[w * kA2WarpDoubles + q];
        dst[e] = a;
    }
}

__global__ void reduce2_kernel(const double* __restrict__ part, const int* __restrict__ work_range, int stride, double* __restrict__ stats) {
    const int g = blockIdx.x;
    const int w0 = work_range[g], w1 = work_range[g + 1];
    for (int e = threadIdx.x; e < stride; e += blockDim.x) {
        double a = 0.0;
        for (int w = w0; w < w1; ++w) a += part[(size_t)w * stride + e];
        stats[(size_t)g * stride + e] = a;
    }
}

// workspace layout of the sorted path, in doubles: [part: n_work_max * stride][ints: chunk_hist | total | base | work | work_range | idx]
struct SortedLayout {
    int n_chunks, n_work_max;
    size_t part_doubles, int_count, total_doubles;
    size_t o_hist, o_total, o_base, o_work, o_range, o_idx;      // offsets in ints
};
static SortedLayout sorted_layout(int64_t total_frames, int n_glob, int dim) {
    SortedLayout L;
    const int stride = 1 + dim + dim * (dim + 1) / 2;
    L.n_chunks = (int)((total_frames + kSortChunk - 1) / kSortChunk);
    if (L.n_chunks < 1) L.n_chunks = 1;
    L.n_work_max = (int)((total_frames + kSplit - 1) / kSplit) + n_glob;
    L.part_doubles = (size_t)L.n_work_max * stride;
    L.o_hist = 0;
    L.o_total = L.o_hist + (size_t)L.n_chunks * n_glob;
    L.o_base = L.o_total + (size_t)n_glob;
    L.o_work = L.o_base + (size_t)n_glob;
    L.o_range = L.o_work + 3 * (size_t)L.n_work_max;
    L.o_idx = L.o_range + (size_t)n_glob + 1;
    L.int_count = L.o_idx + (size_t)total_frames;
    L.total_doubles = L.part_doubles + (L.int_count + 1) / 2;
    return L;
}
static bool sorted_path_ok(int64_t total_frames, int n_glob, int dim) {
    return dim == 39 && n_glob <= kSortMaxGlob && total_frames < (int64_t)0x7fffffff && !getenv("LOE_B200_KMEANS_SCAN");
}

static void kmeans_chunking(int64_t total_frames, int64_t* chunk, int* n_chunks) {
    int64_t ch = 8192;
    while ((total_frames + ch - 1) / ch > 512) ch *= 2;
    *chunk = ch;
    *n_chunks = (int)((total_frames + ch - 1) / ch);
    if (*n_chunks < 1) *n_chunks = 1;
}

}  // namespace loe

extern "C" int loe_align_dev(const int8_t* path_dev, const int64_t* frm_off_dev, int n_utt,
                             const int32_t* tr_off_dev, const int32_t* col_dev, const int32_t* word_dev,
                             const int32_t* word_lo_dev, const int32_t* utt_tr_dev, int remux, int n_glob,
                             uint16_t* bucket_dev, int32_t* counts_dev, void* stream) {
    using namespace loe;
    if (n_utt <= 0) return LOE_OK;
    if (n_glob >= 0xFFFF) { set_error("too many global states"); return LOE_ERR_VALUE; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (!remux && !getenv("LOE_B200_ALIGN_SERIAL")) {
        align_warp_kernel<<<(unsigned)((n_utt + 3) / 4), 128, 0, s>>>(path_dev, frm_off_dev, n_utt, tr_off_dev, col_dev, utt_tr_dev, n_glob,
                                                                     bucket_dev, counts_dev);
        LOE_LAUNCH_CHECK("align_warp_kernel");
        return LOE_OK;
    }
    align_kernel<<<(unsigned)((n_utt + 127) / 128), 128, 0, s>>>(path_dev, frm_off_dev, n_utt, tr_off_dev, col_dev, word_dev,
                                                                word_lo_dev, utt_tr_dev, remux, n_glob, bucket_dev, counts_dev);
    LOE_LAUNCH_CHECK("align_kernel");
    return LOE_OK;
}

extern "C" int64_t loe_kmeans_ws_doubles(int64_t total_frames, int n_glob, int dim) {
    int64_t chunk; int n_chunks;
    loe::kmeans_chunking(total_frames, &chunk, &n_chunks);
    const int64_t scan = (int64_t)n_glob * n_chunks * (1 + dim + dim * (dim + 1) / 2) + n_chunks;    // + per-chunk bucket ranges
    const int64_t sorted = loe::sorted_path_ok(total_frames, n_glob, dim) ? (int64_t)loe::sorted_layout(total_frames, n_glob, dim).total_doubles : 0;
    return scan > sorted ? scan : sorted;
}

extern "C" int loe_kmeans_dev(const float* feat_dev, const uint16_t* bucket_dev, int64_t total_frames, int dim,
                              int n_glob, const float* shift_dev, double* part_ws_dev, double* stats_dev, void* stream) {
    using namespace loe;
    if (n_glob <= 0) return LOE_OK;
    if (dim + 1 > kMaxDimK) { set_error("kmeans kernel supports dim <= %d (got %d)", kMaxDimK - 1, dim); return LOE_ERR_UNSUPPORTED; }
    const int stride = 1 + dim + dim * (dim + 1) / 2;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (total_frames <= 0) { LOE_CUDA(cudaMemsetAsync(stats_dev, 0, sizeof(double) * (size_t)n_glob * stride, s)); return LOE_OK; }
    if (sorted_path_ok(total_frames, n_glob, dim)) {
        const SortedLayout L = sorted_layout(total_frames, n_glob, dim);
        int* ints = reinterpret_cast<int*>(part_ws_dev + L.part_doubles);
        int* hist = ints + L.o_hist; int* total = ints + L.o_total; int* base = ints + L.o_base; int* work = ints + L.o_work; int* range = ints + L.o_range; int* idx = ints + L.o_idx;
        static bool attr_done[64] = {false};
        int dev = 0;
        LOE_CUDA(cudaGetDevice(&dev));
        const size_t scatter_smem = sizeof(int) * 9 * (size_t)n_glob;
        if (dev < 64 && !attr_done[dev]) {
            LOE_CUDA(cudaFuncSetAttribute(bucket_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(int) * 9 * kSortMaxGlob)));
            LOE_CUDA(cudaFuncSetAttribute(accum2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kA2Smem));
            attr_done[dev] = true;
        }
        bucket_hist_kernel<<<(unsigned)L.n_chunks, kSortThreads, sizeof(int) * (size_t)n_glob, s>>>(bucket_dev, total_frames, n_glob, hist);
        LOE_LAUNCH_CHECK("bucket_hist_kernel");
        bucket_scan_kernel<<<(unsigned)((n_glob + 31) / 32), 1024, 0, s>>>(hist, L.n_chunks, n_glob, total);
        LOE_LAUNCH_CHECK("bucket_scan_kernel");
        bucket_plan_kernel<<<1, 1024, 0, s>>>(total, n_glob, base, work, range, L.n_work_max);
        LOE_LAUNCH_CHECK("bucket_plan_kernel");
        bucket_scatter_kernel<<<(unsigned)L.n_chunks, kSortThreads, scatter_smem, s>>>(bucket_dev, total_frames, n_glob, hist, base, idx);
        LOE_LAUNCH_CHECK("bucket_scatter_kernel");
        accum2_kernel<<<(unsigned)L.n_work_max, 256, kA2Smem, s>>>(feat_dev, idx, work, range, n_glob, shift_dev, part_ws_dev);
        LOE_LAUNCH_CHECK("accum2_kernel");
        reduce2_kernel<<<(unsigned)n_glob, 256, 0, s>>>(part_ws_dev, range, stride, stats_dev);
        LOE_LAUNCH_CHECK("reduce2_kernel");
        return LOE_OK;
    }
    int64_t chunk; int n_chunks;
    kmeans_chunking(total_frames, &chunk, &n_chunks);
    const size_t n_part = (size_t)n_glob * n_chunks * stride;
    int* range = reinterpret_cast<int*>(part_ws_dev + n_part);
    LOE_CUDA(cudaMemsetAsync(part_ws_dev, 0, sizeof(double) * n_part, s));       // skipped CTAs leave zeros
    chunk_range_kernel<<<(unsigned)n_chunks, 256, 0, s>>>(bucket_dev, total_frames, chunk, range);
    LOE_LAUNCH_CHECK("chunk_range_kernel");
    dim3 grid((unsigned)n_glob, (unsigned)n_chunks);
    accum_kernel<<<grid, kAccThreads, 0, s>>>(feat_dev, bucket_dev, total_frames, dim, chunk, shift_dev, range, part_ws_dev);
    LOE_LAUNCH_CHECK("accum_kernel");
    reduce_kernel<<<(unsigned)n_glob, 256, 0, s>>>(part_ws_dev, n_chunks, stride, stats_dev);
    LOE_LAUNCH_CHECK("reduce_kernel");
    return LOE_OK;
}
