// Viterbi + backtrace, one CTA per utterance, one thread per trellis position.
//
// Replaces (bit for bit) the three trellis walks of the reference:
//   HiddenMarkovModel._viterbi/_viterbi_static          hidden_markov_model.py:80-91, 160-208
//   HiddenMarkovModelInference._viterbi/_viterbi_static  hidden_markov_model.py:463-581
//   HiddenMarkovModelMultiWord forced alignment          hidden_markov_model.py:591 (same walk as the first,
//                                                        cross-word transitions read back as 0.0)
// Semantics reproduced: float32 candidate = band + delta (one rounding), strict '>' scan from the
// lowest predecessor so the lowest index wins ties, back-pointer 0 when every candidate is -inf
// (np.argmax of an all -inf array), value = float32(max + emission); in the loop grammar the
// word-start positions take max over {penalty + delta[word end w]} (lowest w on ties) and their
// own self loop (wins only if strictly larger), evaluated in float64 when the penalty is an
// np.float64 (the reference default) and in float32 otherwise; termination over END positions
// (lowest index on ties); backtrace with the reference's off-by-one (path[T-1] = s_{T-2}).
//
// State vector: double-buffered in shared memory; back-pointers: uint8 in shared memory
// (global workspace fallback for very long utterances); emission scores are prefetched two
// groups of 8 frames ahead so that the per-frame critical path is shared-memory only.
// Algorithmic HBM bytes per utterance: 4*P*T (scores) + T (path).
#include "common.cuh"
#include "viterbi.cuh"

namespace loe {

constexpr int kPrefetch = 8;


// lowest-index argmax over a warp: (v, i) pairs
template <typename V>
__device__ __forceinline__ void warp_argmax(V& v, int& i) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        V ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}

__global__ void __launch_bounds__(LOE_MAX_POS)
viterbi_kernel(VitArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int u = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t f0 = a.frm_off[u];
    const int T = (int)(a.frm_off[u + 1] - f0);
    if (T <= 0) return;
    const int tr = a.utt_tr ? a.utt_tr[u] : 0;
    const int p0 = a.tr_off[tr];
    const int P = a.tr_off[tr + 1] - p0;

    // shared layout: delta[2][max_pos+2] | ends[max_pos] (int) | path stage | back-pointers
    float* s_delta = reinterpret_cast<float*>(smem_raw);
    const int dstride = a.max_pos + 2;
    int* s_ends = reinterpret_cast<int*>(s_delta + 2 * dstride);
    int* s_nend = s_ends + a.max_pos;
    uint8_t* s_bp = reinterpret_cast<uint8_t*>(s_nend + 4);
    uint8_t* bp = a.bp_in_smem ? s_bp : (a.bp_ws + f0 * LOE_MAX_POS);
    const int bpstride = a.bp_in_smem ? P : LOE_MAX_POS;

    const bool act = tid < P;
    float b0 = neg_inf(), b1 = neg_inf(), b2 = neg_inf();
    int col = 0; unsigned flg = 0;
    if (act) {
        b0 = a.band[(p0 + tid) * 3 + 0]; b1 = a.band[(p0 + tid) * 3 + 1]; b2 = a.band[(p0 + tid) * 3 + 2];
        col = a.col[p0 + tid]; flg = a.flags[p0 + tid];
    }
    const bool is_start = a.loop && (flg & LOE_POS_START);
    // END list in position order (serial scan by thread 0; P <= 128)
    if (tid == 0) {
        int n = 0;
        for (int p = 0; p < P; ++p) if (a.flags[p0 + p] & LOE_POS_END) s_ends[n++] = p;
        *s_nend = n;
    }
    const float* __restrict__ sc = a.scores + f0 * a.ld + col;
    // t = 0
    float e0 = act ? __ldg(sc) : 0.f;
    for (int i = tid; i < 2 * dstride; i += blockDim.x) s_delta[i] = neg_inf();
    __syncthreads();
    if (act && (flg & LOE_POS_INIT)) s_delta[2 + tid] = __fadd_rn(e0, b0);
    const int n_end = *s_nend;
    __syncthreads();

    float ecur[kPrefetch], enext[kPrefetch];
#pragma unroll
    for (int k = 0; k < kPrefetch; ++k) {
        const int t = 1 + k;
        ecur[k] = (act && t < T) ? __ldg(sc + (int64_t)t * a.ld) : 0.f;
    }
    int cur = 0;
    for (int tb = 1; tb < T; tb += kPrefetch) {
#pragma unroll
        for (int k = 0; k < kPrefetch; ++k) {
            const int t = tb + kPrefetch + k;
            enext[k] = (act && t < T) ? __ldg(sc + (int64_t)t * a.ld) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kPrefetch; ++k) {
            const int t = tb + k;
            if (t < T) {                                      // uniform across the CTA
                const float* d = s_delta + cur * dstride + 2;
                float* dn = s_delta + (cur ^ 1) * dstride + 2;
                // cross-word candidate (every warp computes it redundantly: no extra barrier)
                float cross32 = neg_inf(); double cross64 = -CUDART_INF; int cross_arg = 0;
                if (a.loop) {
                    if (a.pen_f64) {
                        double bv = -CUDART_INF; int bi = 0x7fffffff;
                        for (int w0 = 0; w0 < n_end; w0 += 32) {
                            const int w = w0 + lane;
                            double v = (w < n_end) ? a.pen64 + (double)d[s_ends[w]] : -CUDART_INF;
                            int i = (w < n_end) ? w : 0x7fffffff;
                            if (w >= n_end) v = -CUDART_INF;
                            warp_argmax(v, i);
                            if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
                        }
                        if (bi == 0x7fffffff) bi = 0;
                        // np.argmax over NaN-free array: ties -> lowest index; all -inf -> 0
                        cross64 = bv; cross_arg = s_ends[bi];
                    } else {
                        float bv = neg_inf(); int bi = 0x7fffffff;
                        for (int w0 = 0; w0 < n_end; w0 += 32) {
                            const int w = w0 + lane;
                            float v = (w < n_end) ? __fadd_rn(a.pen32, d[s_ends[w]]) : neg_inf();
                            int i = (w < n_end) ? w : 0x7fffffff;
                            warp_argmax(v, i);
                            if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
                        }
                        if (bi == 0x7fffffff) bi = 0;
                        cross32 = bv; cross_arg = s_ends[bi];
                    }
                }
                if (act) {
                    const float e = ecur[k];
                    float val; int arg;
                    if (is_start) {
                        const float selfc = __fadd_rn(b0, d[tid]);
                        if (a.pen_f64) {
                            double mv = cross64; arg = cross_arg;
                            if ((double)selfc > mv) { mv = (double)selfc; arg = tid; }
                            val = __double2float_rn(__dadd_rn(mv, (double)e));
                        } else {
                            float mv = cross32; arg = cross_arg;
                            if (selfc > mv) { mv = selfc; arg = tid; }
                            val = __fadd_rn(mv, e);
                        }
                    } else {
                        float best = __fadd_rn(b2, d[tid - 2]); arg = tid - 2;
                        const float c1 = __fadd_rn(b1, d[tid - 1]);
                        if (c1 > best) { best = c1; arg = tid - 1; }
                        const float c0 = __fadd_rn(b0, d[tid]);
                        if (c0 > best) { best = c0; arg = tid; }
                        if (best == neg_inf()) arg = 0;
                        val = __fadd_rn(best, e);
                    }
                    dn[tid] = val;
                    bp[(int64_t)t * bpstride + tid] = (uint8_t)arg;
                }
                __syncthreads();
                cur ^= 1;
            }
        }
#pragma unroll
        for (int k = 0; k < kPrefetch; ++k) ecur[k] = enext[k];
    }

    // termination + backtrace (thread 0), then a coalesced copy of the path
    const float* d = s_delta + cur * dstride + 2;
    if (a.end_scores) for (int i = tid; i < a.max_ends; i += blockDim.x)
        a.end_scores[(int64_t)u * a.max_ends + i] = (i < n_end) ? d[s_ends[i]] : neg_inf();
    int8_t* s_path = reinterpret_cast<int8_t*>(a.bp_in_smem ? (s_bp + (size_t)a.max_frames * a.max_pos) : s_bp);
    if (tid == 0) {
        float bv = neg_inf(); int bi = 0;
        for (int i = 0; i < n_end; ++i) { const float v = d[s_ends[i]]; if (v > bv) { bv = v; bi = i; } }
        a.best[u] = bi;
        a.best_score[u] = (n_end > 0) ? d[s_ends[bi]] : neg_inf();
        if (T == 1) {
            s_path[0] = -1;
        } else {
            int prev = bp[(int64_t)(T - 1) * bpstride + s_ends[bi]];
            s_path[T - 1] = (int8_t)prev;
            for (int t = T - 2; t >= 0; --t) {
                s_path[t] = (int8_t)prev;
                if (t >= 1) prev = bp[(int64_t)t * bpstride + prev];
            }
        }
    }
    __syncthreads();
    for (int t = tid; t < T; t += blockDim.x) a.path[f0 + t] = s_path[t];
}

// path -> word ids (model_boundary.py:107-147), one thread per utterance
__global__ void labels_kernel(const int8_t* __restrict__ path, const int64_t* __restrict__ frm_off, int n_utt,
                              const int32_t* __restrict__ tr_off, const int32_t* __restrict__ word,
                              const int32_t* __restrict__ word_lo, const int32_t* __restrict__ utt_tr, int skip_label,
                              int8_t* __restrict__ words, int max_words, int32_t* __restrict__ count) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_utt) return;
    const int64_t f0 = frm_off[u];
    const int T = (int)(frm_off[u + 1] - f0);
    const int tr = utt_tr ? utt_tr[u] : 0;
    const int p0 = tr_off[tr];
    const int P = tr_off[tr + 1] - p0;
    int8_t* out = words + (int64_t)u * max_words;
    int n = 0;
    int lo = 0, hi = -1, prev = -1;
    bool bad = T <= 0;
    for (int t = 0; t < T && !bad; ++t) {
        const int cur = path[f0 + t];
        if (cur < 0 || cur >= P) { bad = true; break; }
        if (t > 0 && cur == prev) continue;
        bool emit = false;
        if (t == 0 || cur < lo || cur > hi) {
            lo = word_lo[p0 + cur];
            hi = lo;
            while (hi + 1 < P && word_lo[p0 + hi + 1] == lo) ++hi;
            emit = true;
        } else if (prev == hi && cur == lo) {
            emit = true;
        }
        if (emit) {
            const int lab = word[p0 + cur];
            if (lab != skip_label) { if (n < max_words) out[n] = (int8_t)lab; ++n; }
        }
        prev = cur;
    }
    count[u] = bad ? -1 : n;
}

static size_t vit_smem_bytes(int max_frames, int max_pos, bool bp_in_smem) {
    size_t b = sizeof(float) * 2 * (max_pos + 2) + sizeof(int) * (max_pos + 4);
    if (bp_in_smem) b += (size_t)max_frames * max_pos;
    b += (size_t)max_frames;                 // path stage
    return (b + 15) & ~(size_t)15;
}

constexpr size_t kVitSmemCap = 200 * 1024;

}  // namespace loe

extern "C" int loe_viterbi_bp_fits(int max_frames, int max_pos) {
    return loe::vit_smem_bytes(max_frames, max_pos, true) <= loe::kVitSmemCap ? 1 : 0;
}

extern "C" int loe_viterbi_dev(const float* scores_dev, int ld, const int64_t* frm_off_dev, int n_utt, int max_frames,
                               const int32_t* tr_off_dev, const int32_t* col_dev, const float* band_dev,
                               const uint8_t* flags_dev, int max_pos, const int32_t* utt_tr_dev,
                               int loop, double penalty, int penalty_f64,
                               int8_t* path_dev, float* end_scores_dev, int max_ends,
                               int32_t* best_dev, float* best_score_dev, uint8_t* bp_ws_dev,
                               const int32_t* word_dev, const int32_t* word_lo_dev, int skip_label,
                               int8_t* words_dev, int max_words, int32_t* count_dev, void* stream) {
    using namespace loe;
    if (n_utt <= 0) return LOE_OK;
    if (words_dev && (!word_dev || !word_lo_dev || !count_dev || max_words <= 0)) {
        set_error("label decoding needs word, word_lo, count and max_words > 0");
        return LOE_ERR_VALUE;
    }
    if (max_pos > LOE_MAX_POS) {
        set_error("%d trellis positions: the reference's int8 path/tracer hold at most %d", max_pos, LOE_MAX_POS);
        return LOE_ERR_OVERFLOW;
    }
    if (max_pos <= 0 || max_frames <= 0) { set_error("empty trellis or utterance"); return LOE_ERR_VALUE; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    bool in_smem = vit_smem_bytes(max_frames, max_pos, true) <= kVitSmemCap;
    if (!in_smem) {
        if (!bp_ws_dev) { set_error("back-pointer workspace required for max_frames=%d", max_frames); return LOE_ERR_VALUE; }
        if (vit_smem_bytes(max_frames, max_pos, false) > kVitSmemCap) {
            set_error("utterance of %d frames exceeds the kernel's path staging", max_frames);
            return LOE_ERR_UNSUPPORTED;
        }
    }
    const size_t smem = vit_smem_bytes(max_frames, max_pos, in_smem);
    static bool attr_done[64] = {false};
    int dev = 0;
    LOE_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_done[dev]) {
        LOE_CUDA(cudaFuncSetAttribute(viterbi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVitSmemCap));
        if (dev < 64) attr_done[dev] = true;
    }
    VitArgs a;
    a.scores = scores_dev; a.ld = ld; a.frm_off = frm_off_dev; a.tr_off = tr_off_dev; a.col = col_dev; a.band = band_dev;
    a.flags = flags_dev; a.utt_tr = utt_tr_dev; a.loop = loop; a.pen32 = (float)penalty; a.pen64 = penalty; a.pen_f64 = penalty_f64;
    a.path = path_dev; a.end_scores = end_scores_dev; a.max_ends = max_ends; a.best = best_dev; a.best_score = best_score_dev;
    a.bp_ws = bp_ws_dev; a.bp_in_smem = in_smem ? 1 : 0; a.max_frames = max_frames; a.max_pos = max_pos;
    a.word = word_dev; a.word_lo = word_lo_dev; a.skip_label = skip_label; a.words = words_dev; a.max_words = max_words; a.count = count_dev;
    if (viterbi_warp_launch(a, n_utt, s)) return LOE_OK;      // one warp per utterance (the common case)
    LOE_CUDA(cudaGetLastError());
    const int threads = ((max_pos + 31) / 32) * 32;
    viterbi_kernel<<<(unsigned)n_utt, threads, smem, s>>>(a);
    LOE_LAUNCH_CHECK("viterbi_kernel");
    if (words_dev) {
        labels_kernel<<<(unsigned)((n_utt + 127) / 128), 128, 0, s>>>(path_dev, frm_off_dev, n_utt, tr_off_dev, word_dev, word_lo_dev,
                                                                     utt_tr_dev, skip_label, words_dev, max_words, count_dev);
        LOE_LAUNCH_CHECK("labels_kernel");
    }
    return LOE_OK;
}

extern "C" int loe_labels_dev(const int8_t* path_dev, const int64_t* frm_off_dev, int n_utt,
                              const int32_t* tr_off_dev, const int32_t* word_dev, const int32_t* word_lo_dev,
                              const int32_t* utt_tr_dev, int skip_label,
                              int8_t* words_dev, int max_words, int32_t* count_dev, void* stream) {
    using namespace loe;
    if (n_utt <= 0) return LOE_OK;
    if (max_words <= 0) { set_error("max_words must be positive"); return LOE_ERR_VALUE; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    labels_kernel<<<(unsigned)((n_utt + 127) / 128), 128, 0, s>>>(path_dev, frm_off_dev, n_utt, tr_off_dev, word_dev, word_lo_dev,
                                                                 utt_tr_dev, skip_label, words_dev, max_words, count_dev);
    LOE_LAUNCH_CHECK("labels_kernel");
    return LOE_OK;
}
