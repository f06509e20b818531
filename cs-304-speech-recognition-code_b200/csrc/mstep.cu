// Segmental K-means M-step on the device.
//
// Replaces, for all word models of a batched training run at once, the re-estimation half of
//   HiddenMarkovModelTrainable._update_middleware_parameters   hidden_markov_model.py:320-350
//       means = np.average, convergence test np.allclose(new_means, old_means) BEFORE covariances / transitions are
//       touched (:333-335), np.cov (N-1) + 1e-3 I -> float32, transition counts / row sum -> float32
//   HiddenMarkovModelTrainable._update_inference_weights        :283-292
//       np.log of the transition probabilities, scipy frozen Gaussians (whitening matrix, log pseudo-determinant)
// from the sufficient statistics of kmeans.cu.  The host M-step of round 1 (batched LAPACK eigendecomposition, repacking of
// the tensor-core image, uploads) was twice the device time of an iteration; here nothing leaves the device but one status
// word per model.
//
// Arithmetic that defines the model -- means, convergence test, covariances, transition probabilities -- is done exactly
// as NumPy does it (float64 statistics, the same operation order, one rounding to float32; the np.allclose test in float32
// with atol 1e-8 + rtol 1e-5 |old|), so the float32 parameters are bit-identical to the host M-step's.  What only FEEDS the
// next E-step may differ in the last bits: the whitening matrix is the inverse Cholesky factor of the reversed covariance
// (lower triangular W with W W^T = cov^-1 -- the image emission_h16.cu wants -- instead of scipy's eigenvector form: the
// same quadratic form), log|cov| = 2 sum log(pivot), log-transitions = float32(log(float64(p))).
// A covariance that is not finite, has a Cholesky pivot below 1e-7 of its largest diagonal entry (far above scipy's
// singularity threshold of 2.2e-10 of the largest eigenvalue) or gives |W| >= 2^15 is flagged (LOE_MSTEP_SUSPECT): the host
// then runs scipy's own test on that word, raises what the reference raises, or packs the image itself.
#include <cuda_fp16.h>
#include "common.cuh"

namespace loe {

constexpr int kD = 39;
constexpr int kKW = 40;
constexpr int kStride = 1 + kD + kD * (kD + 1) / 2;       // 820
constexpr int kH16TileHalves = 15 * 240 * 8;             // halves per 6-state tile of the 3xFP16 image (emission_h16.cu)

__device__ __forceinline__ int tri_index(int i, int j) {    // i <= j, row-major upper triangle of a 39 x 39 matrix
    return i * kD - i * (i - 1) / 2 + (j - i);
}

// One CTA per word: new means, the reference's convergence test, bookkeeping of the active set.
__global__ void __launch_bounds__(128)
mstep_mean_kernel(const double* __restrict__ stats, const int32_t* __restrict__ word_first, const int32_t* __restrict__ word_n,
                  float* __restrict__ means32, int32_t* __restrict__ active, int32_t* __restrict__ updated,
                  int32_t* __restrict__ status) {
    const int w = blockIdx.x;
    __shared__ int s_flag[2];                     // [0] some state empty, [1] some mean not close
    if (threadIdx.x == 0) { s_flag[0] = 0; s_flag[1] = 0; }
    __syncthreads();
    if (active[w] != 1) {                         // converged earlier (frozen) or failed: nothing changes
        if (threadIdx.x == 0) { updated[w] = 0; status[w] = 0; }
        return;
    }
    const int g0 = word_first[w], S = word_n[w];
    for (int s = threadIdx.x; s < S; s += blockDim.x)
        if (stats[(size_t)(g0 + s) * kStride] == 0.0) s_flag[0] = 1;
    __syncthreads();
    if (s_flag[0]) {                              // np.concatenate([]) -> HMMTrainMeanFail (:324-331)
        if (threadIdx.x == 0) { updated[w] = 0; status[w] = LOE_MSTEP_MEAN_FAIL; active[w] = -1; }
        return;
    }
    const float atol = 1e-8f, rtol = 1e-5f;
    for (int e = threadIdx.x; e < S * kD; e += blockDim.x) {
        const int g = g0 + e / kD, k = e % kD;
        const double n = stats[(size_t)g * kStride];
        const float old = means32[(size_t)g * kD + k];
        const float nm = (float)((double)old + stats[(size_t)g * kStride + 1 + k] / n);
        // np.isclose in float32: |a - b| <= atol + rtol |b| where b is finite, or a == b
        const bool close = (isfinite(old) && fabsf(__fsub_rn(nm, old)) <= __fadd_rn(atol, __fmul_rn(rtol, fabsf(old)))) || nm == old;
        if (!close) s_flag[1] = 1;
    }
    __syncthreads();
    if (!s_flag[1]) {                             // HMMTrainConverge: the model keeps the previous parameters (:333-335)
        if (threadIdx.x == 0) { updated[w] = 0; status[w] = LOE_MSTEP_CONVERGED; active[w] = 0; }
        return;
    }
    for (int e = threadIdx.x; e < S * kD; e += blockDim.x) {
        const int g = g0 + e / kD, k = e % kD;
        const double n = stats[(size_t)g * kStride];
        const float old = means32[(size_t)g * kD + k];
        means32[(size_t)g * kD + k] = (float)((double)old + stats[(size_t)g * kStride + 1 + k] / n);
    }
    if (threadIdx.x == 0) { updated[w] = 1; status[w] = LOE_MSTEP_UPDATED; }
}

// One CTA per state of an updated word: covariance, whitening matrix + constant into the tensor-core image, the state's
// row of transition probabilities into the trellis band.
__global__ void __launch_bounds__(128)
mstep_cov_kernel(const double* __restrict__ stats, const int32_t* __restrict__ counts, int n_glob,
                 const int32_t* __restrict__ state_word, const int32_t* __restrict__ word_first, const int32_t* __restrict__ word_n,
                 const int32_t* __restrict__ word_tile, const int32_t* __restrict__ updated, const float* __restrict__ means32,
                 float* __restrict__ cov32, int32_t* __restrict__ counts_applied, float* __restrict__ band,
                 __half* __restrict__ b_h16, float* __restrict__ cst_pad, int32_t* __restrict__ status, int32_t* __restrict__ active) {
    const int g = blockIdx.x;
    const int w = state_word[g];
    if (!updated[w]) return;
    const int g0 = word_first[w], S = word_n[w], sl = g - g0;
    const int tid = threadIdx.x;
    __shared__ double a[kD][kKW];                 // reversed covariance -> its Cholesky factor M (lower)
    __shared__ double minv[kD][kKW];              // M^-1 (lower)
    __shared__ double s_mean[kKW];
    __shared__ double s_bias[kKW];
    __shared__ int s_bad;
    if (tid == 0) s_bad = 0;
    const double* st = stats + (size_t)g * kStride;
    const double n = st[0];
    __syncthreads();
    // covariance exactly as the host: (S2 - s1 s1^T / n) / (n - 1) + 1e-3 I, rounded to float32 (:337-343)
    for (int e = tid; e < kD * kD; e += blockDim.x) {
        const int i = e / kD, j = e % kD;
        const int lo = min(i, j), hi = max(i, j);
        const double centred = st[1 + kD + tri_index(lo, hi)] - st[1 + lo] * st[1 + hi] / n;
        const float c32 = (float)(centred / (n - 1.0) + (i == j ? 0.001 : 0.0));
        cov32[(size_t)g * kD * kD + e] = c32;
        if (!isfinite(c32)) s_bad = 1;
        a[kD - 1 - i][kD - 1 - j] = (double)c32;
    }
    if (tid < kD) s_mean[tid] = (double)means32[(size_t)g * kD + tid];
    // transition probabilities of this row and the band entries INTO this state (:344-347, :283-285)
    if (tid == 0) {
        for (int k = 0; k < 3; ++k) {
            const int from = sl - k;
            float lp = -CUDART_INF_F;
            if (from >= 0) {
                long long row = 0;
                for (int j = 0; j < S; ++j) row += counts[(size_t)(g0 + from) * n_glob + g0 + j];
                const float p = (float)((double)counts[(size_t)(g0 + from) * n_glob + g] / (double)row);      // 0 / 0 -> NaN like NumPy
                lp = (float)log((double)p);
            }
            band[(size_t)g * 3 + k] = lp;
        }
    }
    for (int j = tid; j < n_glob; j += blockDim.x) counts_applied[(size_t)g * n_glob + j] = counts[(size_t)g * n_glob + j];
    __syncthreads();
    const bool bad_input = s_bad != 0;
    // right-looking Cholesky of the reversed matrix, in place (lower triangle)
    double log_pdet = 0.0;
    double max_diag = 0.0;
    for (int i = 0; i < kD; ++i) max_diag = fmax(max_diag, fabs(a[i][i]));
    for (int c = 0; c < kD && !bad_input; ++c) {
        const double d = a[c][c];
        if (!(d > 1e-7 * max_diag) || !isfinite(d)) { if (tid == 0) s_bad = 1; break; }     // uniform: every thread reads the same d
        const double l = sqrt(d);
        log_pdet += 2.0 * log(l);
        __syncthreads();
        for (int i = c + tid; i < kD; i += blockDim.x) a[i][c] = (i == c) ? l : a[i][c] / l;
        __syncthreads();
        const int m = kD - 1 - c;                 // trailing block (c+1 .. 38)^2, lower triangle
        for (int e = tid; e < m * m; e += blockDim.x) {
            const int i = c + 1 + e / m, j = c + 1 + e % m;
            if (j <= i) a[i][j] -= a[i][c] * a[j][c];
        }
        __syncthreads();
    }
    __syncthreads();
    if (s_bad) {                                  // no image for this state: the word is parked until the host has looked at it
        if (tid == 0) { atomicOr(status + w, LOE_MSTEP_SUSPECT); active[w] = -2; }
        return;
    }
    // M^-1 by forward substitution, one column per thread
    if (tid < kD) {
        const int j = tid;
        for (int i = 0; i < kD; ++i) minv[i][j] = 0.0;
        for (int i = j; i < kD; ++i) {
            double acc = (i == j) ? 1.0 : 0.0;
            for (int k = j; k < i; ++k) acc -= a[i][k] * minv[k][j];
            minv[i][j] = acc / a[i][i];
        }
    }
    __syncthreads();
    // W[k][j] = M^-1[38 - j][38 - k] (lower triangular); bias row 39 = -mean . W; column 39 = 0
    if (tid < kKW) {
        const int j = tid;
        double b = 0.0;
        if (j < kD)
            for (int k = j; k < kD; ++k) b -= s_mean[k] * minv[kD - 1 - j][kD - 1 - k];
        s_bias[j] = b;
    }
    __syncthreads();
    const int tile = word_tile[w] + sl / 6, slot = sl % 6;
    __half* img = b_h16 + (size_t)tile * kH16TileHalves;
    int bad = 0;
    for (int e = tid; e < kKW * kKW; e += blockDim.x) {
        const int k = e / kKW, j = e % kKW;
        double v = 0.0;
        if (j < kD) v = (k < kD) ? ((j <= k) ? minv[kD - 1 - j][kD - 1 - k] : 0.0) : s_bias[j];
        if (!(fabs(v) < 32768.0)) bad = 1;
        const __half hi = __double2half(v);
        const __half lo = __double2half(v - (double)__half2float(hi));
        const int kc = k >> 3, q = k & 7, nn = (j >> 3) * 48 + slot * 8 + (j & 7);
        img[((size_t)kc * 240 + nn) * 8 + q] = hi;
        img[((size_t)(5 + kc) * 240 + nn) * 8 + q] = lo;
        img[((size_t)(10 + kc) * 240 + nn) * 8 + q] = hi;
    }
    if (bad) { atomicOr(status + w, LOE_MSTEP_SUSPECT); active[w] = -2; }
    if (tid == 0) cst_pad[tile * 6 + slot] = (float)(-0.5 * ((double)kD * 1.8378770664093453 + log_pdet));
}

}  // namespace loe

extern "C" int loe_mstep_dev(const double* stats_dev, const int32_t* counts_dev, int n_glob, int n_words,
                             const int32_t* state_word_dev, const int32_t* word_first_dev, const int32_t* word_n_dev,
                             const int32_t* word_tile_dev, float* means32_dev, float* cov32_dev, int32_t* counts_applied_dev,
                             float* band_dev, void* b_h16_dev, float* cst_pad_dev, int32_t* active_dev, int32_t* updated_dev,
                             int32_t* status_dev, int dim, void* stream) {
    using namespace loe;
    if (n_words <= 0 || n_glob <= 0) return LOE_OK;
    if (dim != kD) { set_error("the device M-step is built for dim == 39 (got %d)", dim); return LOE_ERR_UNSUPPORTED; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    mstep_mean_kernel<<<(unsigned)n_words, 128, 0, s>>>(stats_dev, word_first_dev, word_n_dev, means32_dev, active_dev, updated_dev, status_dev);
    LOE_LAUNCH_CHECK("mstep_mean_kernel");
    mstep_cov_kernel<<<(unsigned)n_glob, 128, 0, s>>>(stats_dev, counts_dev, n_glob, state_word_dev, word_first_dev, word_n_dev, word_tile_dev,
                                                     updated_dev, means32_dev, cov32_dev, counts_applied_dev, band_dev,
                                                     static_cast<__half*>(b_h16_dev), cst_pad_dev, status_dev, active_dev);
    LOE_LAUNCH_CHECK("mstep_cov_kernel");
    return LOE_OK;
}
