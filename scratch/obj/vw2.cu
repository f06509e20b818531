// Viterbi + backtrace, ONE WARP per utterance (same semantics as viterbi.cu, see there for the
// reference lines).  Each lane owns SPL consecutive trellis positions (SPL = 1, 2 or 4 for up to
// 32 / 64 / 128 positions), the state vector lives in registers, the three predecessors come from
// two warp shuffles, and there is no block barrier in the time loop.
//
// Word-start rule of the loop grammar without a value/index shuffle reduction: float addition is
// monotonic, so  max_w fl(pen + d_w) = fl(pen + max_w d_w);  the maximum of the word-end scores is
// one float REDUX, and np.argmax's "lowest index among equal candidates" is the lowest set bit of a
// ballot of  fl(pen + d_w) == max.
//
// Back-pointers are 2-bit codes (0/1/2 = came from p, p-1, p-2; 3 = word-start took the cross-word
// candidate recorded per frame, or "all candidates -inf -> position 0"), 16 codes per word,
// so a 460-frame, 58-state utterance needs 7.4 KB of shared memory instead of 26.7 KB and four
// utterances share a CTA.  A word holds the codes of ONE position for 16 consecutive frames: the
// backtrace then moves in RUNS -- one shared-memory load and one find-leading-one give the next
// frame at which the path leaves its position (self loops dominate: ~70 runs for 400 frames), and
// the whole warp writes the run's stretch of the path with one store.
#include "viterbi.cuh"
#include <type_traits>

namespace loe {

constexpr int kWarpsPerCta = 2;
constexpr int kPre = 8;
constexpr size_t kWarpSmemCap = 200 * 1024;

// warp-wide float maximum in one instruction (sm_100a: CREDUX.MAX.F32); the scores are never NaN
__device__ __forceinline__ float warp_max(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}

struct WarpLayout {
    int n_rows;          // back-pointer rows of 32 words
    int off_cross, off_path, off_ends, off_flags, off_wlo, off_whi, off_wlab, off_endp, total;
};

__host__ __device__ inline WarpLayout warp_layout(int max_frames, int max_pos, int spl) {
    WarpLayout L;
    L.n_rows = ((max_frames + 15) / 16) * spl;        // [16-frame group][slot]: 32 words, one per lane
    int o = L.n_rows * 128;
    L.off_cross = o; o += max_frames * 4 + 4;            // one word per frame: ballot of the lanes reaching the cross-word maximum
    L.off_path = o;  o += (max_frames + 3) & ~3;
    L.off_ends = o;  o += max_pos * 4;
    L.off_flags = o; o += (max_pos + 3) & ~3;
    L.off_wlo = o;   o += (max_pos + 3) & ~3;
    L.off_whi = o;   o += (max_pos + 3) & ~3;
    L.off_wlab = o;  o += (max_pos + 3) & ~3;
    L.off_endp = o;  o += 32;
    L.total = (o + 15) & ~15;
    return L;
}

template <int SPL, bool LOOP, bool PENF64>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
viterbi_warp_kernel(VitArgs a, int n_utt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int FPW = 16;                              // frames per back-pointer word (one position)
    static_assert(kPre == 8, "two blocks of the time loop fill a back-pointer word");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = blockIdx.x * kWarpsPerCta + warp;
    if (u >= n_utt) return;
    const WarpLayout L = warp_layout(a.max_frames, a.max_pos, SPL);
    unsigned char* base = smem_raw + (size_t)warp * L.total;
    uint32_t* s_bp = reinterpret_cast<uint32_t*>(base);
    uint32_t* s_cross = reinterpret_cast<uint32_t*>(base + L.off_cross);   // per frame: ballot (or position) of the cross-word argmax
    int8_t* s_path = reinterpret_cast<int8_t*>(base + L.off_path);
    uint8_t* s_flags = base + L.off_flags;

    const int64_t f0 = a.frm_off[u];
    const int T = (int)(a.frm_off[u + 1] - f0);
    if (T <= 0) return;
    const int tr = a.utt_tr ? a.utt_tr[u] : 0;
    const int p0 = a.tr_off[tr];
    const int P = a.tr_off[tr + 1] - p0;

    float b0[SPL], b1[SPL], b2[SPL], d[SPL];
    bool act[SPL], is_start[SPL], is_end[SPL];
    const float* __restrict__ src[SPL];                  // running pointer into the score matrix
    const float* __restrict__ sc = a.scores + f0 * a.ld;
    unsigned end_mask[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        const int p = lane * SPL + i;
        act[i] = p < P;
        b0[i] = b1[i] = b2[i] = neg_inf();
        unsigned flg = 0; int col = 0;
        if (act[i]) {
            b0[i] = a.band[(p0 + p) * 3 + 0]; b1[i] = a.band[(p0 + p) * 3 + 1]; b2[i] = a.band[(p0 + p) * 3 + 2];
            col = a.col[p0 + p]; flg = a.flags[p0 + p];
            s_flags[p] = (uint8_t)flg;
        }
        src[i] = sc + col;
        is_start[i] = LOOP && (flg & LOE_POS_START);
        is_end[i] = (flg & LOE_POS_END) != 0;
        end_mask[i] = __ballot_sync(FULL, is_end[i]);
    }
    // Usual case (every word model has at least SPL states): no lane owns two END positions.  The cross-word argmax
    // then needs one candidate per lane -- one add, one compare, one ballot -- and records the winning LANE; the
    // backtrace maps it to the END position through a 32-entry table.
    int my_ends = 0, end_slot = 0;
#pragma unroll
    for (int i = 0; i < SPL; ++i) { if (is_end[i]) { ++my_ends; end_slot = i; } }
    const bool one_end = LOOP && !__any_sync(FULL, my_ends > 1);
    const bool has_end = my_ends == 1;
    uint8_t* s_endp = base + L.off_endp;
    s_endp[lane] = (uint8_t)(has_end ? lane * SPL + end_slot : 0);
    // positions 0 and 1 have no p - 1 / p - 2: their bands are -inf (the trellis builder writes them so; enforced here
    // because the time loop does not mask the shuffles of lane 0)
    if (lane == 0) {
        b1[0] = neg_inf(); b2[0] = neg_inf();
        if (SPL >= 2) b2[SPL >= 2 ? 1 : 0] = neg_inf();
    }
    if (SPL == 1 && lane == 1) b2[0] = neg_inf();
    // END positions are ordered by position = (lane, slot) lexicographically
    int n_end = 0, lower = 0;
#pragma unroll
    for (int i = 0; i < SPL; ++i) { lower += __popc(end_mask[i] & ((1u << lane) - 1)); n_end += __popc(end_mask[i]); }

    // max over END positions of d, and the lowest END position p with pred(d_p) true
    auto end_max = [&]() -> float {
        float lm = neg_inf();
#pragma unroll
        for (int i = 0; i < SPL; ++i) lm = is_end[i] ? fmaxf(lm, d[i]) : lm;
        return warp_max(lm);
    };

    // t = 0
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        const float e0 = __ldg(src[i]);
        const unsigned flg = act[i] ? s_flags[lane * SPL + i] : 0u;
        d[i] = (act[i] && (flg & LOE_POS_INIT)) ? __fadd_rn(e0, b0[i]) : neg_inf();
        src[i] += a.ld;
    }

    float ecur[kPre][SPL], enext[kPre][SPL];
#pragma unroll
    for (int k = 0; k < kPre; ++k)
#pragma unroll
        for (int i = 0; i < SPL; ++i)
            ecur[k][i] = (1 + k < T) ? __ldg(src[i] + (int64_t)k * a.ld) : 0.f;
    uint32_t bits[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) bits[i] = 0;
    // frame t = 1 + j;  j runs in blocks of kPre (jb is a multiple of kPre)
    for (int jb = 0; jb < T - 1; jb += kPre) {
#pragma unroll
        for (int i = 0; i < SPL; ++i) src[i] += (int64_t)kPre * a.ld;
        if (jb + 2 * kPre < T) {                            // the whole next block exists: no tests
#pragma unroll
            for (int k = 0; k < kPre; ++k)
#pragma unroll
                for (int i = 0; i < SPL; ++i) enext[k][i] = __ldg(src[i] + (int64_t)k * a.ld);
        } else {
#pragma unroll
            for (int k = 0; k < kPre; ++k)
#pragma unroll
                for (int i = 0; i < SPL; ++i)
                    enext[k][i] = (1 + jb + kPre + k < T) ? __ldg(src[i] + (int64_t)k * a.ld) : 0.f;
        }
        uint32_t blk[SPL];                                  // the codes of this block: 2 bits per frame
#pragma unroll
        for (int i = 0; i < SPL; ++i) blk[i] = 0;
        // One frame of the recursion.  FAST = the block lies strictly before the utterance's last frame: no bounds test per
        // frame and the partial back-pointer / cross-word words are only flushed at their static positions.
        auto frame = [&](auto fast_tag, const int k) {
            constexpr bool FAST = decltype(fast_tag)::value;
            const int j = jb + k;
            if (!FAST && !(j < T - 1)) return;
            // ---- cross-word candidate: fl(pen + max END d); argmax = lowest END position reaching it
            float cross32 = neg_inf(); double cross64 = -CUDART_INF; int cross_arg = 0;
            if (LOOP) {
                if (one_end) {
                    float lm = d[0];
#pragma unroll
                    for (int i = 1; i < SPL; ++i) lm = (end_slot == i) ? d[i] : lm;
                    lm = has_end ? lm : neg_inf();
                    const float m = warp_max(lm);
                    unsigned eq;
                    if (PENF64) {
                        cross64 = __dadd_rn(a.pen64, (double)m);
                        eq = __ballot_sync(FULL, has_end && __dadd_rn(a.pen64, (double)lm) == cross64);
                    } else {
                        cross32 = __fadd_rn(a.pen32, m);
                        eq = __ballot_sync(FULL, has_end && __fadd_rn(a.pen32, lm) == cross32);
                    }
                    cross_arg = (int)eq;                         // the lanes reaching the maximum; the backtrace takes the lowest
                } else {
                    const float m = end_max();
                    int pos = 0x7fffffff;
                    if (PENF64) {
                        cross64 = __dadd_rn(a.pen64, (double)m);
#pragma unroll
                        for (int i = 0; i < SPL; ++i) {
                            const unsigned eq = __ballot_sync(FULL, __dadd_rn(a.pen64, (double)d[i]) == cross64) & end_mask[i];
                            if (eq) pos = min(pos, (__ffs(eq) - 1) * SPL + i);
                        }
                    } else {
                        cross32 = __fadd_rn(a.pen32, m);
#pragma unroll
                        for (int i = 0; i < SPL; ++i) {
                            const unsigned eq = __ballot_sync(FULL, __fadd_rn(a.pen32, d[i]) == cross32) & end_mask[i];
                            if (eq) pos = min(pos, (__ffs(eq) - 1) * SPL + i);
                        }
                    }
                    cross_arg = (pos == 0x7fffffff) ? 0 : pos;
                }
                if (lane == 0) s_cross[j] = (uint32_t)cross_arg;
            }
            // ---- predecessors held by the lane below
            float up1 = __shfl_up_sync(FULL, d[SPL - 1], 1);
            float up2 = (SPL >= 2) ? __shfl_up_sync(FULL, d[SPL >= 2 ? SPL - 2 : 0], 1) : __shfl_up_sync(FULL, d[0], 2);
            // (lane 0 -- and lane 1 for SPL = 1 -- receive their own values: the bands of the positions that have no
            // p - 1 / p - 2 were set to -inf above, so those candidates are -inf whatever the shuffle delivers)
            float nd[SPL];
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                const float p1 = (i >= 1) ? d[i >= 1 ? i - 1 : 0] : up1;
                const float p2 = (i >= 2) ? d[i >= 2 ? i - 2 : 0] : (i == 1 ? up1 : up2);
                const float e = ecur[k][i];
                // candidates in the reference's order p-2, p-1, p; a later one only wins when strictly larger
                const float c2 = __fadd_rn(b2[i], p2), c1 = __fadd_rn(b1[i], p1), c0 = __fadd_rn(b0[i], d[i]);
                const float best = fmaxf(fmaxf(c2, c1), c0);
                unsigned code = (c2 == best) ? 2u : (c1 == best) ? 1u : 0u;
                if (best == neg_inf()) code = 3;
                float val = __fadd_rn(best, e);
                if (LOOP) {
                    // word start: b1 = b2 = -inf, so best == c0 (its self loop); the cross-word
                    // candidate wins unless the self loop is strictly larger
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
                }
                nd[i] = val;                 // positions beyond the trellis have -inf bands and stay at -inf by themselves
                blk[i] |= code << (2 * k);
            }
#pragma unroll
            for (int i = 0; i < SPL; ++i) d[i] = nd[i];
        };
        const bool fast = jb + kPre < T - 1;
        if (fast) {
#pragma unroll
            for (int k = 0; k < kPre; ++k) frame(std::true_type{}, k);
        } else {
#pragma unroll
            for (int k = 0; k < kPre; ++k) frame(std::false_type{}, k);
        }
        // the block's codes go into the low or the high half of the 16-frame words; a word is stored when it is full
        // and after the utterance's last block
        const int half = jb & kPre;
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            bits[i] |= blk[i] << (2 * half);
            if (half || !fast) { s_bp[((jb >> 4) * SPL + i) * 32 + lane] = bits[i]; bits[i] = 0; }
        }
#pragma unroll
        for (int k = 0; k < kPre; ++k)
#pragma unroll
            for (int i = 0; i < SPL; ++i) ecur[k][i] = enext[k][i];
    }
    __syncwarp();

    // ---- termination: best END (lowest position on ties); END scores out
    const float m = end_max();
    int pos = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        const unsigned eq = __ballot_sync(FULL, is_end[i] && d[i] == m);
        if (eq) pos = min(pos, (__ffs(eq) - 1) * SPL + i);
    }
    const int end_pos = (pos == 0x7fffffff) ? 0 : pos;
    int bi = 0;                                          // rank of end_pos among END positions
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        const int ln = end_pos / SPL, sl = end_pos % SPL;
        const unsigned below = (i < sl) ? ((ln == 31) ? 0xffffffffu : ((2u << ln) - 1)) : ((1u << ln) - 1);
        bi += __popc(end_mask[i] & below);
    }
    if (lane == 0) { a.best[u] = bi; a.best_score[u] = (n_end > 0) ? m : neg_inf(); }
    if (a.end_scores) {
        int mine = 0;
#pragma unroll
        for (int i = 0; i < SPL; ++i)
            if (is_end[i]) { a.end_scores[(int64_t)u * a.max_ends + lower + mine] = d[i]; ++mine; }
        for (int w = n_end + lane; w < a.max_ends; w += 32) a.end_scores[(int64_t)u * a.max_ends + w] = neg_inf();
    }

    // ---- backtrace in runs (every lane walks the same chain).  q_t = position at frame t, q_(T-1) = end_pos; the code
    // of the step into frame t is slot (t - 1) & 15 of word [(t - 1) >> 4][q_t].  The reference records the
    // predecessor one frame late at the end (path[T-1] = path[T-2] = q_(T-2)).
    if (T == 1) {
        if (lane == 0) s_path[0] = -1;
    } else {
        int pcur = n_end > 0 ? end_pos : 0;
        int t_hi = T - 1;                               // frames (t_lo, t_hi] found so far hold position pcur
        while (t_hi >= 1) {
            const int j = t_hi - 1, g = j >> 4, sl = j & 15;
            const uint32_t w = s_bp[(g * SPL + pcur % SPL) * 32 + pcur / SPL];
            const uint32_t nz = (w | (w >> 1)) & 0x55555555u & ((2u << (2 * sl)) - 1u);   // steps <= j of this word that leave pcur
            const int k = nz ? (31 - __clz(nz)) >> 1 : -1;
            const int t_lo = (g << 4) + k + 1;          // frames t_lo .. t_hi hold pcur (k = -1: down to the word's first frame's predecessor)
            if (t_lo + lane <= t_hi) s_path[t_lo + lane] = (int8_t)pcur;      // a run lies inside one word: at most 17 frames
            if (nz) {
                const unsigned code = (w >> (2 * k)) & 3u;
                const int jj = (g << 4) + k;
                int pn = pcur - (int)code;
                if (code == 3) {
                    pn = 0;
                    if (LOOP && (s_flags[pcur] & LOE_POS_START)) {
                        const uint32_t cw = s_cross[jj];
                        pn = one_end ? (cw ? (int)s_endp[__ffs(cw) - 1] : 0) : (int)cw;
                    }
                }
                pcur = pn;
            }
            t_hi = nz ? t_lo - 1 : t_lo;                // no step left in this word: frame t_lo holds pcur too, go on in the word before
        }
        if (lane == 0) s_path[0] = (int8_t)pcur;
        __syncwarp();
        if (lane == 0) s_path[T - 1] = s_path[T - 2];
    }
    __syncwarp();
    for (int t = lane; t < T; t += 32) a.path[f0 + t] = s_path[t];

    // ---- fused label decoding (model_boundary.py:107-147).  The reference's running "current word
    // range" is always the range of the word that holds the previous state, so whether a change point
    // emits a word is a local test:  different word, or same word re-entered first-state-from-last-state.
    if (a.words) {
        uint8_t* s_wlo = base + L.off_wlo;
        uint8_t* s_whi = base + L.off_whi;
        uint8_t* s_wlab = base + L.off_wlab;
        for (int p = lane; p < P; p += 32) {
            const int lo = a.word_lo[p0 + p];
            int hi = p;
            while (hi + 1 < P && a.word_lo[p0 + hi + 1] == lo) ++hi;
            s_wlo[p] = (uint8_t)lo; s_whi[p] = (uint8_t)hi; s_wlab[p] = (uint8_t)a.word[p0 + p];
        }
        __syncwarp();
        int8_t* out = a.words + (int64_t)u * a.max_words;
        int n = 0;
        bool bad = false;
        for (int tb = 0; tb < T; tb += 32) {
            const int t = tb + lane;
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
            const unsigned m = __ballot_sync(FULL, emit);
            if (emit) { const int k = n + __popc(m & ((1u << lane) - 1)); if (k < a.max_words) out[k] = (int8_t)lab; }
            n += __popc(m);
        }
        bad = __any_sync(FULL, bad);
        if (lane == 0) a.count[u] = bad ? -1 : n;
    }
}

template <int SPL, bool LOOP, bool PENF64>
static bool launch(const VitArgs& a, int n_utt, cudaStream_t s) {
    const WarpLayout L = warp_layout(a.max_frames, a.max_pos, SPL);
    const size_t smem = (size_t)L.total * kWarpsPerCta;
    if (smem > kWarpSmemCap) return false;
    static bool attr_done[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (dev >= 64 || !attr_done[dev]) {
        if (cudaFuncSetAttribute(viterbi_warp_kernel<SPL, LOOP, PENF64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)kWarpSmemCap) != cudaSuccess)
            return false;
        if (dev < 64) attr_done[dev] = true;
    }
    viterbi_warp_kernel<SPL, LOOP, PENF64><<<(unsigned)((n_utt + kWarpsPerCta - 1) / kWarpsPerCta), kWarpsPerCta * 32, smem, s>>>(a, n_utt);
    return cudaGetLastError() == cudaSuccess;
}

template <int SPL>
static bool launch_spl(const VitArgs& a, int n_utt, cudaStream_t s) {
    if (!a.loop) return launch<SPL, false, false>(a, n_utt, s);
    return a.pen_f64 ? launch<SPL, true, true>(a, n_utt, s) : launch<SPL, true, false>(a, n_utt, s);
}

bool viterbi_warp_launch(const VitArgs& a, int n_utt, cudaStream_t s) {
    if (a.max_pos <= 32) return launch_spl<1>(a, n_utt, s);
    if (a.max_pos <= 64) return launch_spl<2>(a, n_utt, s);
    return launch_spl<4>(a, n_utt, s);
}

}  // namespace loe
