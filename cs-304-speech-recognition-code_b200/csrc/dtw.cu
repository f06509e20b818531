// Time-synchronous template DTW.  Replaces DynamicTimeWarping.search
// (dynamic_time_wrapping.py:66-116) for a batch of samples against one template set, bit for bit:
// float32 local distance sqrt(sum((a-b)^2)) in NumPy's pairwise summation order, float64 accumulated
// costs, predecessors {insertion (i, j-1), shrink (i-2, j-1), match (i-1, j-1)} with the shrink step
// confined to the word, beam pruning against (1 + factor) x the previous column's minimum, and the
// reference's quirks: the row shared by the end of word w-1 and the start of word w is evaluated
// twice per column (the start-of-word value survives, both feed the column minimum, the first one's
// path code survives when the second is pruned), row 0 wraps to the last template frame / last cost
// row, and word w's distance is read one row above its last frame.
//
// One CTA per sample: the two live columns of the cost matrix are float64 arrays in shared memory,
// thread t owns rows t, t+256, ...; one barrier pair per column (values, then the column minimum).
#include "common.cuh"

namespace loe {

constexpr int kDtwThreads = 256;
constexpr int kDtwMaxDim = 64;

__device__ __forceinline__ float np_sum_le128(const float* a, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    }
    float r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], a[i + k]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
}

struct DtwArgs {
    const float* seq; int H; int D;
    const int32_t* row_start;      // [H+1] first row of the word whose frames include row i (as its LAST-word view)
    const int32_t* row_is_boundary;// [H+1] 1 if row i is start_w of a word w > 0
    const int32_t* starts; const int32_t* lens; int W;
    const float* samp; const int64_t* samp_off;
    int pruning; double factor;
    double* dist; int32_t* best_idx; double* best_dist;
    double* cost_out; int8_t* path_out;       // optional, sample 0 only
};

template <int DT>                      // DT > 0: compile-time feature dimension (distance loop in registers)
__global__ void __launch_bounds__(kDtwThreads)
dtw_kernel(DtwArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* col0 = reinterpret_cast<double*>(smem_raw);
    double* col1 = col0 + (a.H + 1);
    float* s_x = reinterpret_cast<float*>(col1 + (a.H + 1));      // current sample frame [D]
    __shared__ double s_red[kDtwThreads / 32];
    __shared__ double s_min;
    const int u = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t f0 = a.samp_off[u];
    const int L = (int)(a.samp_off[u + 1] - f0);
    const int H = a.H, D = (DT > 0) ? DT : a.D;
    const double INF = CUDART_INF;
    const bool dump = (u == 0) && a.cost_out != nullptr;

    // column 0: zero at every word start, inf elsewhere
    for (int i = tid; i <= H; i += kDtwThreads) {
        const bool is_start = (i == 0) || (a.row_is_boundary[i] != 0);
        col0[i] = is_start ? 0.0 : INF;
        if (dump) { a.cost_out[(int64_t)i * (L + 1)] = col0[i]; if (a.path_out) a.path_out[(int64_t)i * (L + 1)] = 0; }
    }
    if (tid == 0) s_min = INF;
    __syncthreads();
    double* prev = col0;
    double* cur = col1;
    for (int j = 1; j <= L; ++j) {
        if (tid < D) s_x[tid] = a.samp[(f0 + j - 1) * D + tid];
        __syncthreads();
        const double thr = s_min * (1.0 + a.factor);               // previous column's minimum (inf for j = 1)
        double local_min = INF;
        for (int i = tid; i <= H; i += kDtwThreads) {
            // local distance, NumPy order
            const float* __restrict__ sr = a.seq + (int64_t)((i == 0) ? (H - 1) : (i - 1)) * D;
            float sq[DT > 0 ? DT : kDtwMaxDim];
            if (DT > 0) {
#pragma unroll
                for (int k = 0; k < DT; ++k) { const float df = __fsub_rn(__ldg(sr + k), s_x[k]); sq[k] = __fmul_rn(df, df); }
            } else {
#pragma unroll 1
                for (int k = 0; k < D; ++k) { const float df = __fsub_rn(__ldg(sr + k), s_x[k]); sq[k] = __fmul_rn(df, df); }
            }
            const double d = (double)__fsqrt_rn(np_sum_le128(sq, D));
            const double ins = prev[i];
            const double mat = prev[(i == 0) ? H : (i - 1)];
            const bool boundary = a.row_is_boundary[i] != 0;
            double value = INF; int code = 0;
            // version A: row i as a frame row of the word that ENDS here or contains it (skipped for i = 0,
            // which only exists as the start row of word 0)
            if (i > 0) {
                const int st = a.row_start[i];
                const double shr = (i - 2 < st) ? INF : prev[i - 2];
                const double m = fmin(ins, fmin(shr, mat));
                const double c = d + m;
                if (!(a.pruning && c > thr)) {
                    value = c;
                    code = (m == ins) ? 1 : ((m == shr) ? 2 : 3);
                    if (c != INF) local_min = fmin(local_min, c);
                }
            }
            // version B: row i as the start row of word w (i = 0, or a boundary): overwrites version A
            if (i == 0 || boundary) {
                const double m = fmin(ins, mat);                   // shrink is outside the word: inf
                const double c = d + m;
                if (!(a.pruning && c > thr)) {
                    value = c;
                    code = (m == ins) ? 1 : ((m == INF) ? 2 : 3);  // min == shrink(inf) only when everything is inf
                    if (c != INF) local_min = fmin(local_min, c);
                } else {
                    value = INF;                                   // pruned: cost inf, path code of version A survives
                }
            }
            cur[i] = value;
            if (dump) { a.cost_out[(int64_t)i * (L + 1) + j] = value; if (a.path_out) a.path_out[(int64_t)i * (L + 1) + j] = (int8_t)code; }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) local_min = fmin(local_min, __shfl_xor_sync(0xffffffffu, local_min, o));
        if (lane == 0) s_red[warp] = local_min;
        __syncthreads();
        if (tid == 0) {
            double m = INF;
            for (int w = 0; w < kDtwThreads / 32; ++w) m = fmin(m, s_red[w]);
            s_min = m;
        }
        double* t = prev; prev = cur; cur = t;
        __syncthreads();
    }
    // distances: one row above each word's last frame (:106-107), then first minimum
    for (int w = tid; w < a.W; w += kDtwThreads) a.dist[(int64_t)u * a.W + w] = prev[a.starts[w] + a.lens[w] - 1];
    __syncthreads();
    if (tid == 0) {
        double best = INF; int bi = 0;
        for (int w = 0; w < a.W; ++w) {
            const double v = prev[a.starts[w] + a.lens[w] - 1];
            if (w == 0 || v < best) { best = v; bi = w; }
        }
        a.best_idx[u] = bi; a.best_dist[u] = best;
    }
}

}  // namespace loe

extern "C" int loe_dtw_dev(const float* seq_dev, int n_rows, int dim, const int32_t* row_start_dev,
                           const int32_t* row_is_boundary_dev, const int32_t* starts_dev, const int32_t* lens_dev,
                           int n_words, const float* samp_dev, const int64_t* samp_off_dev, int n_samples,
                           int pruning, double pruning_factor, double* dist_dev, int32_t* best_idx_dev,
                           double* best_dist_dev, double* cost_out_dev, int8_t* path_out_dev, void* stream) {
    using namespace loe;
    if (n_samples <= 0) return LOE_OK;
    if (dim <= 0 || dim > kDtwMaxDim) { set_error("DTW kernel supports 1 <= dim <= %d (got %d)", kDtwMaxDim, dim); return LOE_ERR_UNSUPPORTED; }
    if (n_rows <= 0 || n_words <= 0) { set_error("empty template set"); return LOE_ERR_VALUE; }
    const size_t smem = sizeof(double) * 2 * (size_t)(n_rows + 1) + sizeof(float) * kDtwMaxDim;
    if (smem > 200 * 1024) { set_error("template set of %d frames exceeds the kernel's shared-memory columns", n_rows); return LOE_ERR_UNSUPPORTED; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    auto kern = (dim == 39) ? dtw_kernel<39> : dtw_kernel<0>;
    LOE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DtwArgs a;
    a.seq = seq_dev; a.H = n_rows; a.D = dim; a.row_start = row_start_dev; a.row_is_boundary = row_is_boundary_dev;
    a.starts = starts_dev; a.lens = lens_dev; a.W = n_words; a.samp = samp_dev; a.samp_off = samp_off_dev;
    a.pruning = pruning; a.factor = pruning_factor; a.dist = dist_dev; a.best_idx = best_idx_dev; a.best_dist = best_dist_dev;
    a.cost_out = cost_out_dev; a.path_out = path_out_dev;
    kern<<<(unsigned)n_samples, kDtwThreads, smem, s>>>(a);
    LOE_LAUNCH_CHECK("dtw_kernel");
    return LOE_OK;
}
