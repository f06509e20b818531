// Gaussian emission scoring, SIMT path (float32 and float64).  Replaces
// MultivariateNormal.log_pdf (hidden_markov_model.py:46-48 -> scipy _logpdf):
//     score[f][s] = cst[s] - 0.5 * | (x_f - mean_s) . U_s |^2
//
// One thread owns one frame: its feature row lives in registers for the whole kernel, the
// whitening matrix U_s of the current state is staged in shared memory (double buffered) and broadcast to every thread with 128-bit loads, the 39 whitened coordinates are
// register accumulators, and the [frames x states] tile of results goes through shared memory
// so that the global store is row-contiguous.  This is the parity anchor and the fallback for
// dimensions the tensor-core kernel (emission_tc.cu) does not cover.
#include "common.cuh"

namespace loe {

constexpr int kFramesPerBlockE = 128;

template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

// DP = dim padded to a multiple of 4.  U staged as [dim][DP] (+ mean [DP]) per state.
template <typename T, int DIM>
__global__ void __launch_bounds__(kFramesPerBlockE)
emission_simt_kernel(const float* __restrict__ feat, int64_t n_frames, const T* __restrict__ mean,
                     const T* __restrict__ U, const T* __restrict__ cst, int n_states,
                     float* __restrict__ out, int ld_out) {
    constexpr int DP = (DIM + 3) & ~3;
    constexpr int kStateChunk = 16;                     // result tile width staged in smem
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_u = reinterpret_cast<T*>(smem_raw);            // [2][DIM*DP + DP]
    float* s_res = reinterpret_cast<float*>(s_u + 2 * (DIM * DP + DP));   // [128][kStateChunk+1]
    constexpr int kStage = DIM * DP + DP;

    const int tid = threadIdx.x;
    const int64_t fbase = (int64_t)blockIdx.x * kFramesPerBlockE;
    const int64_t f = fbase + tid;
    const bool live = f < n_frames;

    // feature row -> registers.  The tile is contiguous in global memory, so it is read with
    // consecutive lanes on consecutive floats and transposed through shared memory (the U
    // buffers, not yet in use), 32 frames at a time.
    T x[DIM];
    {
        const int n_in_block = (int)min((int64_t)kFramesPerBlockE, n_frames - fbase);
        const float* __restrict__ src = feat + fbase * DIM;
        float* s_t = reinterpret_cast<float*>(s_u);      // [32][DIM+1]
        const int total = n_in_block * DIM;
        for (int r0 = 0; r0 < kFramesPerBlockE; r0 += 32) {
            const int lo = r0 * DIM, hi = min(total, (r0 + 32) * DIM);
            for (int i = lo + tid; i < hi; i += kFramesPerBlockE) {
                const int rr = i / DIM - r0, cc = i % DIM;
                s_t[rr * (DIM + 1) + cc] = src[i];
            }
            __syncthreads();
            if (tid >= r0 && tid < r0 + 32) {
#pragma unroll
                for (int i = 0; i < DIM; ++i) x[i] = live ? (T)s_t[(tid - r0) * (DIM + 1) + i] : (T)0;
            }
            __syncthreads();
        }
    }

    auto stage_load = [&](int s, int buf) {
        // U_s rows padded to DP; mean appended.  Plain loads (tiny, L2 resident).
        T* dst = s_u + buf * kStage;
        const T* us = U + (size_t)s * DIM * DIM;
        for (int i = tid; i < DIM * DP; i += kFramesPerBlockE) {
            const int r = i / DP, c = i - r * DP;
            dst[i] = (c < DIM) ? us[r * DIM + c] : (T)0;
        }
        for (int i = tid; i < DP; i += kFramesPerBlockE) dst[DIM * DP + i] = (i < DIM) ? mean[(size_t)s * DIM + i] : (T)0;
    };

    stage_load(0, 0);
    __syncthreads();
    for (int s = 0; s < n_states; ++s) {
        const int buf = s & 1;
        if (s + 1 < n_states) stage_load(s + 1, buf ^ 1);
        const T* us = s_u + buf * kStage;
        const T* mu = us + DIM * DP;
        T y[DP];
#pragma unroll
        for (int j = 0; j < DP; ++j) y[j] = (T)0;
#pragma unroll
        for (int i = 0; i < DIM; ++i) {
            const T d = x[i] - mu[i];
#pragma unroll
            for (int j = 0; j < DP; j += 4) {
                const typename Vec4<T>::type w = *reinterpret_cast<const typename Vec4<T>::type*>(us + i * DP + j);
                y[j + 0] = fma(d, w.x, y[j + 0]);
                y[j + 1] = fma(d, w.y, y[j + 1]);
                y[j + 2] = fma(d, w.z, y[j + 2]);
                y[j + 3] = fma(d, w.w, y[j + 3]);
            }
        }
        T maha = (T)0;
#pragma unroll
        for (int j = 0; j < DIM; ++j) maha = fma(y[j], y[j], maha);
        const int sc = s % kStateChunk;
        s_res[tid * (kStateChunk + 1) + sc] = (float)(cst[s] - (T)0.5 * maha);
        __syncthreads();                                  // next stage landed; result column visible
        if (sc == kStateChunk - 1 || s == n_states - 1) {
            const int s_lo = s - sc, w = sc + 1;
            const int64_t n_in_block = min((int64_t)kFramesPerBlockE, n_frames - fbase);
            for (int i = tid; i < (int)n_in_block * w; i += kFramesPerBlockE) {
                const int r = i / w, c = i - r * w;
                out[(fbase + r) * ld_out + s_lo + c] = s_res[r * (kStateChunk + 1) + c];
            }
            __syncthreads();
        }
    }
}

template <typename T, int DIM>
static int launch_simt(const float* feat, int64_t n_frames, const void* mean, const void* U, const void* cst,
                       int n_states, float* out, int ld_out, cudaStream_t s) {
    constexpr int DP = (DIM + 3) & ~3;
    const size_t smem = sizeof(T) * 2 * (DIM * DP + DP) + sizeof(float) * kFramesPerBlockE * 17;
    static_assert(sizeof(T) * 2 * (DIM * DP + DP) >= sizeof(float) * 32 * (DIM + 1), "feature staging needs the U buffers");
    auto kern = emission_simt_kernel<T, DIM>;
    LOE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((n_frames + kFramesPerBlockE - 1) / kFramesPerBlockE);
    kern<<<grid, kFramesPerBlockE, smem, s>>>(feat, n_frames, (const T*)mean, (const T*)U, (const T*)cst, n_states, out, ld_out);
    LOE_LAUNCH_CHECK("emission_simt_kernel");
    return LOE_OK;
}

}  // namespace loe

extern "C" int loe_emission_dev(const float* feat_dev, int64_t n_frames, int dim,
                                const void* mean_dev, const void* u_dev, const void* cst_dev, int n_states,
                                float* out_dev, int ld_out, int precision, void* stream) {
    using namespace loe;
    if (n_frames <= 0 || n_states <= 0) return LOE_OK;
    if (ld_out < n_states) { set_error("ld_out (%d) < n_states (%d)", ld_out, n_states); return LOE_ERR_VALUE; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (precision != 0 && precision != 1) {
        set_error("unknown precision %d (the tensor-core path is loe_emission_tc_dev)", precision);
        return LOE_ERR_VALUE;
    }
#define LOE_DISPATCH_DIM(D)                                                                                     \
    case D:                                                                                                     \
        return precision == 0 ? launch_simt<float, D>(feat_dev, n_frames, mean_dev, u_dev, cst_dev, n_states, out_dev, ld_out, s) \
                              : launch_simt<double, D>(feat_dev, n_frames, mean_dev, u_dev, cst_dev, n_states, out_dev, ld_out, s);
    switch (dim) {
        LOE_DISPATCH_DIM(39)
        LOE_DISPATCH_DIM(13)
        LOE_DISPATCH_DIM(26)
        default:
            set_error("emission kernels are built for dim in {13, 26, 39} (got %d)", dim);
            return LOE_ERR_UNSUPPORTED;
    }
#undef LOE_DISPATCH_DIM
}
