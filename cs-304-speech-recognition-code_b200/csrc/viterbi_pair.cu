// Loop-grammar Viterbi + backtrace + labels with TWO utterances per warp (same semantics as viterbi_warp.cu / viterbi.cu,
// see there for the reference lines: hidden_markov_model.py:463-581, model_boundary.py:107-147).
//
// The digit-loop trellis of the reference's decoder has 58 positions.  One warp per utterance (viterbi_warp.cu, 2 positions
// per lane) spends about half of its ~145 instructions per frame on work that does not depend on the number of positions a
// lane owns: the REDUX of the word-end maximum, the ballots of the lowest-index argmax, the two shuffles that fetch the
// predecessors from the lane below, the packing and flushing of back-pointers.  Here a HALF-warp owns an utterance, 4
// positions per lane (15 lanes x 4 = 60 >= 58): those per-frame instructions are issued once for two utterances (ballots and
// shuffles are warp-wide anyway; the REDUX runs on the two half masks), so the warp scheduler sees roughly half the
// instructions per utterance-frame.  Used for loop decoding with 33..60 positions; everything else stays with the
// one-warp kernel.
//
// Frames past the end of the shorter utterance of a pair are computed but not committed (the two halves must stay in
// lock step for the warp-wide intrinsics).  Back-pointers: 2-bit codes, one 32-bit word per lane and 4 frames.
#include "viterbi.cuh"
#include <stdlib.h>
#include <type_traits>

namespace loe {

constexpr int kPairWarps = 4;              // warps per CTA: 8 utterances
constexpr int kPairPre = 8;                // frames of scores in registers ahead of the recursion
constexpr int kPairSpl = 4;
constexpr size_t kPairSmemCap = 200 * 1024;

__device__ __forceinline__ int pkey(float f) {
    const int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float pkey_inv(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

struct PairLayout {
    int n_rows;          // back-pointer rows of 16 words (one per lane of the half-warp), 4 frames each
    int off_cross, off_path, off_flags, off_wlo, off_whi, off_wlab, total;
};

__host__ __device__ inline PairLayout pair_layout(int max_frames, int max_pos) {
    PairLayout L;
    L.n_rows = (max_frames + 3) / 4;
    int o = L.n_rows * 64;
    L.off_cross = o; o += ((max_frames + 3) & ~3) + 4;
    L.off_path = o;  o += (max_frames + 3) & ~3;
    L.off_flags = o; o += (max_pos + 3) & ~3;
    L.off_wlo = o;   o += (max_pos + 3) & ~3;
    L.off_whi = o;   o += (max_pos + 3) & ~3;
    L.off_wlab = o;  o += (max_pos + 3) & ~3;
    L.total = (o + 15) & ~15;
    return L;
}

template <bool PENF64>
__global__ void __launch_bounds__(kPairWarps * 32)
viterbi_pair_kernel(VitArgs a, int n_utt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int SPL = kPairSpl;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = lane >> 4, hl = lane & 15, sh = 16 * half;
    const unsigned hmask = half ? 0xffff0000u : 0x0000ffffu;
    const int u0 = (blockIdx.x * kPairWarps + warp) * 2;
    if (u0 >= n_utt) return;                                   // whole warp
    const int u = u0 + half;
    const bool have = u < n_utt;                               // the last warp of an odd batch has an empty second half
    const PairLayout L = pair_layout(a.max_frames, a.max_pos);
    unsigned char* base = smem_raw + (size_t)(warp * 2 + half) * L.total;
    uint32_t* s_bp = reinterpret_cast<uint32_t*>(base);
    uint32_t* s_cross = reinterpret_cast<uint32_t*>(base + L.off_cross);   // 4 frames per word
    int8_t* s_path = reinterpret_cast<int8_t*>(base + L.off_path);
    uint8_t* s_flags = base + L.off_flags;

    const int64_t f0 = have ? a.frm_off[u] : 0;
    const int T = have ? (int)(a.frm_off[u + 1] - f0) : 0;
    const int tr = (have && a.utt_tr) ? a.utt_tr[u] : 0;
    const int p0 = a.tr_off[tr];
    const int P = have ? a.tr_off[tr + 1] - p0 : 0;
    const int Tw = max(T, __shfl_xor_sync(FULL, T, 16));       // frames of the longer utterance of the pair
    const int Tmin = min(T, __shfl_xor_sync(FULL, T, 16));

    float b0[SPL], b1[SPL], b2[SPL], d[SPL];
    bool act[SPL], is_start[SPL], is_end[SPL];
    const float* __restrict__ src[SPL];
    const float* __restrict__ sc = a.scores + f0 * a.ld;
    unsigned end_mask[SPL];                                    // END positions of slot i, one bit per lane of THIS half
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        const int p = hl * SPL + i;
        act[i] = p < P;
        b0[i] = b1[i] = b2[i] = neg_inf();
        unsigned flg = 0; int col = 0;
        if (act[i]) {
            b0[i] = a.band[(p0 + p) * 3 + 0]; b1[i] = a.band[(p0 + p) * 3 + 1]; b2[i] = a.band[(p0 + p) * 3 + 2];
            col = a.col[p0 + p]; flg = a.flags[p0 + p];
            s_flags[p] = (uint8_t)flg;
        }
        src[i] = sc + col;
        is_start[i] = (flg & LOE_POS_START) != 0;
        is_end[i] = (flg & LOE_POS_END) != 0;
        end_mask[i] = (__ballot_sync(FULL, is_end[i]) >> sh) & 0xffffu;
    }
    int n_end = 0, lower = 0;
#pragma unroll
    for (int i = 0; i < SPL; ++i) { lower += __popc(end_mask[i] & ((1u << hl) - 1)); n_end += __popc(end_mask[i]); }

    auto end_max = [&]() -> float {
        float lm = neg_inf();
#pragma unroll
        for (int i = 0; i < SPL; ++i) lm = is_end[i] ? fmaxf(lm, d[i]) : lm;
        return pkey_inv(__reduce_max_sync(hmask, pkey(lm)));
    };

    // t = 0
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        const float e0 = (act[i] && T > 0) ? __ldg(src[i]) : 0.f;
        const unsigned flg = act[i] ? s_flags[hl * SPL + i] : 0u;
        d[i] = (act[i] && (flg & LOE_POS_INIT)) ? __fadd_rn(e0, b0[i]) : neg_inf();
        src[i] += a.ld;
    }

    float ecur[kPairPre][SPL], enext[kPairPre][SPL];
#pragma unroll
    for (int k = 0; k < kPairPre; ++k)
#pragma unroll
        for (int i = 0; i < SPL; ++i)
            ecur[k][i] = (act[i] && 1 + k < T) ? __ldg(src[i] + (int64_t)k * a.ld) : 0.f;
    uint32_t bits = 0, cbits = 0;
    // frame t = 1 + j;  j runs in blocks of kPairPre
    for (int jb = 0; jb < Tw - 1; jb += kPairPre) {
#pragma unroll
        for (int i = 0; i < SPL; ++i) src[i] += (int64_t)kPairPre * a.ld;
#pragma unroll
        for (int k = 0; k < kPairPre; ++k)
#pragma unroll
            for (int i = 0; i < SPL; ++i)
                enext[k][i] = (act[i] && 1 + jb + kPairPre + k < T) ? __ldg(src[i] + (int64_t)k * a.ld) : 0.f;
        // One frame of the recursion for both utterances.  FAST = the block lies strictly before the last frame of BOTH:
        // no liveness tests, back-pointer / cross-word words flushed at their static positions only.
        auto frame = [&](auto fast_tag, const int k) {
            constexpr bool FAST = decltype(fast_tag)::value;
            const int j = jb + k;
            const bool live = FAST || j < T - 1;                // this half still has a frame t = j + 1
            // ---- cross-word candidate: fl(pen + max END d); argmax = lowest END position reaching it
            const float m = end_max();
            int pos = 0x7fffffff;
            float cross32 = neg_inf(); double cross64 = -CUDART_INF;
            if (PENF64) {
                cross64 = __dadd_rn(a.pen64, (double)m);
#pragma unroll
                for (int i = 0; i < SPL; ++i) {
                    const unsigned eq = (__ballot_sync(FULL, __dadd_rn(a.pen64, (double)d[i]) == cross64) >> sh) & end_mask[i];
                    if (eq) pos = min(pos, (__ffs(eq) - 1) * SPL + i);
                }
            } else {
                cross32 = __fadd_rn(a.pen32, m);
#pragma unroll
                for (int i = 0; i < SPL; ++i) {
                    const unsigned eq = (__ballot_sync(FULL, __fadd_rn(a.pen32, d[i]) == cross32) >> sh) & end_mask[i];
                    if (eq) pos = min(pos, (__ffs(eq) - 1) * SPL + i);
                }
            }
            const int cross_arg = (pos == 0x7fffffff) ? 0 : pos;
            cbits |= (uint32_t)cross_arg << (8 * (k & 3));
            if (live && ((k & 3) == 3 || (!FAST && j == T - 2))) s_cross[j >> 2] = cbits;      // every lane of the half: same value
            if ((k & 3) == 3) cbits = 0;
            // ---- predecessors held by the lane below (inside the half)
            float up1 = __shfl_up_sync(FULL, d[SPL - 1], 1, 16);
            float up2 = __shfl_up_sync(FULL, d[SPL - 2], 1, 16);
            if (hl == 0) { up1 = neg_inf(); up2 = neg_inf(); }
            float nd[SPL];
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                const float p1 = (i >= 1) ? d[i >= 1 ? i - 1 : 0] : up1;
                const float p2 = (i >= 2) ? d[i >= 2 ? i - 2 : 0] : (i == 1 ? up1 : up2);
                const float e = ecur[k][i];
                float best = __fadd_rn(b2[i], p2); unsigned code = 2;
                const float c1 = __fadd_rn(b1[i], p1);
                if (c1 > best) { best = c1; code = 1; }
                const float c0 = __fadd_rn(b0[i], d[i]);
                if (c0 > best) { best = c0; code = 0; }
                if (best == neg_inf()) code = 3;
                float val = __fadd_rn(best, e);
                // word start: b1 = b2 = -inf, so best == c0 (its self loop); the cross-word candidate wins unless the self
                // loop is strictly larger
                if (PENF64) {
                    const bool use_cross = is_start[i] && !((double)c0 > cross64);
                    const float vx = __double2float_rn(__dadd_rn(cross64, (double)e));
                    val = use_cross ? vx : val;
                    code = use_cross ? 3u : code;
                } else {
                    const bool use_cross = is_start[i] && !(c0 > cross32);
                    const float vx = __fadd_rn(cross32, e);
                    val = use_cross ? vx : val;
                    code = use_cross ? 3u : code;
                }
                nd[i] = val;
                bits |= code << (2 * ((k & 3) * SPL + i));
            }
#pragma unroll
            for (int i = 0; i < SPL; ++i) d[i] = live ? nd[i] : d[i];
            if (live && ((k & 3) == 3 || (!FAST && j == T - 2))) s_bp[(j >> 2) * 16 + hl] = bits;
            if ((k & 3) == 3) bits = 0;
        };
        if (jb + kPairPre < Tmin - 1) {
#pragma unroll
            for (int k = 0; k < kPairPre; ++k) frame(std::true_type{}, k);
        } else {
#pragma unroll
            for (int k = 0; k < kPairPre; ++k) frame(std::false_type{}, k);
        }
#pragma unroll
        for (int k = 0; k < kPairPre; ++k)
#pragma unroll
            for (int i = 0; i < SPL; ++i) ecur[k][i] = enext[k][i];
    }
    __syncwarp();

    // ---- termination: best END (lowest position on ties); END scores out
    const float m = end_max();
    int pos = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        const unsigned eq = (__ballot_sync(FULL, is_end[i] && d[i] == m) >> sh) & 0xffffu;
        if (eq) pos = min(pos, (__ffs(eq) - 1) * SPL + i);
    }
    const int end_pos = (pos == 0x7fffffff) ? 0 : pos;
    int bi = 0;                                          // rank of end_pos among END positions
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        const int ln = end_pos / SPL, sl = end_pos % SPL;
        const unsigned below = (i < sl) ? ((2u << ln) - 1) : ((1u << ln) - 1);
        bi += __popc(end_mask[i] & below);
    }
    if (have && hl == 0) { a.best[u] = bi; a.best_score[u] = (n_end > 0) ? m : neg_inf(); }
    if (have && a.end_scores) {
        int mine = 0;
#pragma unroll
        for (int i = 0; i < SPL; ++i)
            if (is_end[i]) { a.end_scores[(int64_t)u * a.max_ends + lower + mine] = d[i]; ++mine; }
        for (int w = n_end + hl; w < a.max_ends; w += 16) a.end_scores[(int64_t)u * a.max_ends + w] = neg_inf();
    }

    // ---- backtrace (every lane of the half walks the same chain; its lane 0 records it).  frame t >= 1 is j = t - 1
    auto decode = [&](int t, int p) -> int {
        const int j = t - 1;
        const uint32_t w = s_bp[(j >> 2) * 16 + p / SPL];
        const unsigned code = (w >> (2 * ((j & 3) * SPL + p % SPL))) & 3u;
        if (code < 3) return p - (int)code;
        if (s_flags[p] & LOE_POS_START) return (int)((s_cross[j >> 2] >> (8 * (j & 3))) & 0xffu);
        return 0;
    };
    __syncwarp();
    if (T == 1) {
        if (hl == 0) s_path[0] = -1;
    } else if (T > 1) {
        int prev = decode(T - 1, n_end > 0 ? end_pos : 0);
        if (hl == 0) s_path[T - 1] = (int8_t)prev;
        for (int t = T - 2; t >= 0; --t) {
            if (hl == 0) s_path[t] = (int8_t)prev;
            if (t >= 1) prev = decode(t, prev);
        }
    }
    __syncwarp();
    for (int t = hl; t < T; t += 16) a.path[f0 + t] = s_path[t];

    // ---- fused label decoding (model_boundary.py:107-147), 16 frames of a half per round
    if (a.words) {
        uint8_t* s_wlo = base + L.off_wlo;
        uint8_t* s_whi = base + L.off_whi;
        uint8_t* s_wlab = base + L.off_wlab;
        for (int p = hl; p < P; p += 16) {
            const int lo = a.word_lo[p0 + p];
            int hi = p;
            while (hi + 1 < P && a.word_lo[p0 + hi + 1] == lo) ++hi;
            s_wlo[p] = (uint8_t)lo; s_whi[p] = (uint8_t)hi; s_wlab[p] = (uint8_t)a.word[p0 + p];
        }
        __syncwarp();
        int8_t* out = a.words + (int64_t)(have ? u : 0) * a.max_words;
        int n = 0;
        bool bad = false;
        for (int tb = 0; tb < Tw; tb += 16) {
            const int t = tb + hl;
            bool emit = false; int lab = 0;
            if (t < T) {
                const int cur = s_path[t];
                if (cur < 0 || cur >= P) bad = true;
                else {
                    lab = s_wlab[cur];
                    if (t == 0) emit = true;
                    else {
                        const int prev = s_path[t - 1];
                        if (prev >= 0 && prev < P && cur != prev)
                            emit = (s_wlo[cur] != s_wlo[prev]) || (prev == s_whi[prev] && cur == s_wlo[cur]);
                    }
                }
            }
            emit = emit && lab != a.skip_label;
            const unsigned mk = (__ballot_sync(FULL, emit) >> sh) & 0xffffu;
            if (emit) { const int k = n + __popc(mk & ((1u << hl) - 1)); if (k < a.max_words) out[k] = (int8_t)lab; }
            n += __popc(mk);
        }
        const unsigned any_bad = (__ballot_sync(FULL, bad) >> sh) & 0xffffu;
        if (have && hl == 0) a.count[u] = any_bad ? -1 : n;
    }
}

template <bool PENF64>
static bool pair_launch(const VitArgs& a, int n_utt, cudaStream_t s) {
    const PairLayout L = pair_layout(a.max_frames, a.max_pos);
    const size_t smem = (size_t)L.total * kPairWarps * 2;
    if (smem > kPairSmemCap) return false;
    static bool attr_done[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (dev >= 64 || !attr_done[dev]) {
        if (cudaFuncSetAttribute(viterbi_pair_kernel<PENF64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPairSmemCap) != cudaSuccess)
            return false;
        if (dev < 64) attr_done[dev] = true;
    }
    const int per_cta = kPairWarps * 2;
    viterbi_pair_kernel<PENF64><<<(unsigned)((n_utt + per_cta - 1) / per_cta), kPairWarps * 32, smem, s>>>(a, n_utt);
    return cudaGetLastError() == cudaSuccess;
}

// Loop decoding with 33..60 trellis positions and more than one utterance: two utterances per warp.  Returns false when
// the shape does not apply (the caller falls through to the one-warp kernel).
bool viterbi_pair_launch(const VitArgs& a, int n_utt, cudaStream_t s) {
    if (!a.loop || a.max_pos <= 32 || a.max_pos > 15 * kPairSpl || n_utt < 2 || getenv("LOE_B200_VITERBI_NO_PAIR")) return false;
    return a.pen_f64 ? pair_launch<true>(a, n_utt, s) : pair_launch<false>(a, n_utt, s);
}

}  // namespace loe
