// Energy-hysteresis silence stripper.  Replaces SignalSeparation._remove_empty /
// detect_speech (signal_separation.py:103-164) for a batch of utterances, bit for bit:
//   - per-frame energy = float32 mean |x| with NumPy's pairwise summation order (np.average on a
//     float32 frame: 8 interleaved accumulators per <=128-element block, blocks split at n/2 rounded
//     down to a multiple of 8), so every threshold decision matches the reference,
//   - thresholds = high/low * max|x| evaluated in float64 like the reference's np.float64 scalars,
//   - the hysteresis state machine with its quirks: the frame that trips the silence counter is
//     credited to the noise list but not to the result; the trailing partial frame (possibly empty:
//     mean of nothing = NaN = "no speech") takes part.
// One CTA per utterance: block max-reduce, energies from shared-memory staged tiles (row stride
// frame_size + 1: conflict free), then one thread walks the frames.
#include "common.cuh"

namespace loe {

constexpr int kVadThreads = 128;
constexpr int kVadTile = 64;                 // frames staged per tile

__device__ __forceinline__ float np_block_sum(const float* a, int n) {      // n <= 128
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

template <int DEPTH>
__device__ __forceinline__ float np_pairwise_sum(const float* a, int n) {
    if (n <= 128) return np_block_sum(a, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_pairwise_sum<DEPTH - 1>(a, n2), np_pairwise_sum<DEPTH - 1>(a + n2, n - n2));
}
template <>
__device__ __forceinline__ float np_pairwise_sum<0>(const float* a, int n) { return np_block_sum(a, n < 128 ? n : 128); }

template <typename SampleT>
__global__ void __launch_bounds__(kVadThreads)
silence_kernel(const SampleT* __restrict__ pcm, const int64_t* __restrict__ pcm_off, int frame_size,
               double high, double low, int max_silence_frames, const int64_t* __restrict__ efrm_off,
               float* __restrict__ energy, uint8_t* __restrict__ noise, int32_t* __restrict__ seg,
               float* __restrict__ max_out) {
    extern __shared__ __align__(16) float s_tile[];              // [kVadTile][frame_size + 1]
    __shared__ float s_red[kVadThreads / 32];
    const int u = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t s0 = pcm_off[u];
    const int64_t n = pcm_off[u + 1] - s0;
    const SampleT* __restrict__ x = pcm + s0;
    const int64_t e0 = efrm_off[u];
    const int n_frames = (int)(efrm_off[u + 1] - e0);            // n / frame_size full frames + 1 partial
    const int n_full = (int)(n / frame_size);

    // ---- max |x|
    float m = 0.f;
    for (int64_t i = tid; i < n; i += kVadThreads) m = fmaxf(m, fabsf((float)x[i]));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    m = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    if (tid == 0) max_out[u] = m;
    const double hi_thr = __dmul_rn(high, (double)m), lo_thr = __dmul_rn(low, (double)m);

    // ---- energies, kVadTile frames at a time
    const int stride = frame_size + 1;
    for (int fb = 0; fb < n_frames; fb += kVadTile) {
        const int nf = min(kVadTile, n_frames - fb);
        const int64_t base = (int64_t)fb * frame_size;
        const int64_t cnt = min((int64_t)nf * frame_size, n - base);
        __syncthreads();
        for (int64_t i = tid; i < cnt; i += kVadThreads) {
            const int r = (int)(i / frame_size), c = (int)(i - (int64_t)r * frame_size);
            s_tile[r * stride + c] = fabsf((float)x[base + i]);
        }
        __syncthreads();
        if (tid < nf) {
            const int f = fb + tid;
            const int len = (f < n_full) ? frame_size : (int)(n - (int64_t)n_full * frame_size);
            const float sum = np_pairwise_sum<3>(s_tile + tid * stride, len);
            energy[e0 + f] = __fdiv_rn(sum, (float)len);             // 0/0 = NaN for the empty trailing frame
        }
    }
    __syncthreads();
    __threadfence_block();

    // ---- hysteresis state machine (signal_separation.py:118-146)
    if (tid == 0) {
        bool between = false, ever = false, done = false;
        int counter = 0, start = -1, end = n_frames;
        for (int f = 0; f < n_frames; ++f) {
            const double e = (double)energy[e0 + f];
            bool tripped = false, is_noise = false;
            if (between) {
                if (e > lo_thr) counter = 0;
                else { between = false; ++counter; tripped = counter >= max_silence_frames; }
            } else {
                if (e > hi_thr) { between = ever = true; counter = 0; if (start < 0) start = f; }
                else { is_noise = true; if (ever) { ++counter; tripped = counter >= max_silence_frames; } }
            }
            noise[e0 + f] = is_noise ? 1 : 0;
            if (tripped) { done = true; end = f; for (int g = f + 1; g < n_frames; ++g) noise[e0 + g] = 0; break; }
        }
        if (start < 0) start = end;
        seg[u * 4 + 0] = done ? 1 : 0; seg[u * 4 + 1] = start; seg[u * 4 + 2] = end; seg[u * 4 + 3] = n_frames;
    }
}

}  // namespace loe

extern "C" int loe_silence_dev(const void* pcm_dev, int pcm_format, const int64_t* pcm_off_dev, int n_utt,
                               int frame_size, double high, double low, int max_silence_frames,
                               const int64_t* efrm_off_dev, float* energy_dev, uint8_t* noise_dev, int32_t* seg_dev,
                               float* max_dev, void* stream) {
    using namespace loe;
    if (n_utt <= 0) return LOE_OK;
    if (frame_size <= 0 || frame_size > 1024) { set_error("frame_size %d outside [1, 1024]", frame_size); return LOE_ERR_VALUE; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = sizeof(float) * kVadTile * (frame_size + 1);
    if (pcm_format == LOE_PCM_F32) {
        LOE_CUDA(cudaFuncSetAttribute(silence_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        silence_kernel<float><<<(unsigned)n_utt, kVadThreads, smem, s>>>((const float*)pcm_dev, pcm_off_dev, frame_size, high, low,
                                                                        max_silence_frames, efrm_off_dev, energy_dev, noise_dev, seg_dev, max_dev);
    } else if (pcm_format == LOE_PCM_S16) {
        LOE_CUDA(cudaFuncSetAttribute(silence_kernel<short>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        silence_kernel<short><<<(unsigned)n_utt, kVadThreads, smem, s>>>((const short*)pcm_dev, pcm_off_dev, frame_size, high, low,
                                                                        max_silence_frames, efrm_off_dev, energy_dev, noise_dev, seg_dev, max_dev);
    } else { set_error("unknown pcm_format %d", pcm_format); return LOE_ERR_VALUE; }
    LOE_LAUNCH_CHECK("silence_kernel");
    return LOE_OK;
}
