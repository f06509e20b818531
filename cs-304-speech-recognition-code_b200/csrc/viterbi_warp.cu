// Viterbi + backtrace, ONE WARP per utterance (same semantics as viterbi.cu, see there for the
// reference lines).  Each lane owns SPL consecutive trellis positions (SPL = 1, 2 or 4 for up to
// 32 / 64 / 128 positions), the state vector lives in registers, the three predecessors come from
// two warp shuffles, and there is no block barrier in the time loop.
//
// Word-start rule of the loop grammar without a value/index shuffle reduction: float addition is
// monotonic, so  max_w fl(pen + d_w) = fl(pen + max_w d_w);  the maximum of the word-end scores is
// one REDUX on an order-preserving integer key, and np.argmax's "lowest index among equal
// candidates" is the lowest set bit of a ballot of  fl(pen + d_w) == max.
//
// Back-pointers are 2-bit codes (0/1/2 = came from p, p-1, p-2; 3 = word-start took the cross-word
// candidate recorded per frame, or "all candidates -inf -> position 0"), 16 codes per lane word,
// so a 460-frame, 58-state utterance needs 7.4 KB of shared memory instead of 26.7 KB and four
// utterances share a CTA.
#include "viterbi.cuh"

namespace loe {

constexpr int kWarpsPerCta = 4;
constexpr int kPre = 8;
constexpr size_t kWarpSmemCap = 200 * 1024;

__device__ __forceinline__ int fkey(float f) {
    const int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float fkey_inv(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

struct WarpLayout {
    int n_rows;          // back-pointer rows of 32 words
    int off_cross, off_path, off_ends, off_flags, total;
};

__host__ __device__ inline WarpLayout warp_layout(int max_frames, int max_pos, int spl) {
    WarpLayout L;
    const int fpw = 16 / spl;
    L.n_rows = (max_frames + fpw - 1) / fpw;
    int o = L.n_rows * 128;
    L.off_cross = o; o += (max_frames + 3) & ~3;
    L.off_path = o;  o += (max_frames + 3) & ~3;
    L.off_ends = o;  o += max_pos * 4;
    L.off_flags = o; o += (max_pos + 3) & ~3;
    L.total = (o + 15) & ~15;
    return L;
}

template <int SPL>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
viterbi_warp_kernel(VitArgs a, int n_utt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int FPW = 16 / SPL;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = blockIdx.x * kWarpsPerCta + warp;
    if (u >= n_utt) return;
    const WarpLayout L = warp_layout(a.max_frames, a.max_pos, SPL);
    unsigned char* base = smem_raw + (size_t)warp * L.total;
    uint32_t* s_bp = reinterpret_cast<uint32_t*>(base);
    uint8_t* s_cross = base + L.off_cross;
    int8_t* s_path = reinterpret_cast<int8_t*>(base + L.off_path);
    int* s_ends = reinterpret_cast<int*>(base + L.off_ends);
    uint8_t* s_flags = base + L.off_flags;

    const int64_t f0 = a.frm_off[u];
    const int T = (int)(a.frm_off[u + 1] - f0);
    if (T <= 0) return;
    const int tr = a.utt_tr ? a.utt_tr[u] : 0;
    const int p0 = a.tr_off[tr];
    const int P = a.tr_off[tr + 1] - p0;

    float b0[SPL], b1[SPL], b2[SPL], d[SPL];
    int col[SPL];
    unsigned flg[SPL];
    bool act[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        const int p = lane * SPL + i;
        act[i] = p < P;
        b0[i] = b1[i] = b2[i] = neg_inf(); col[i] = 0; flg[i] = 0;
        if (act[i]) {
            b0[i] = a.band[(p0 + p) * 3 + 0]; b1[i] = a.band[(p0 + p) * 3 + 1]; b2[i] = a.band[(p0 + p) * 3 + 2];
            col[i] = a.col[p0 + p]; flg[i] = a.flags[p0 + p];
            s_flags[p] = (uint8_t)flg[i];
        }
    }
    // END positions in order: rank by ballot prefix
    int n_end = 0;
    {
        unsigned m[SPL];
#pragma unroll
        for (int i = 0; i < SPL; ++i) m[i] = __ballot_sync(FULL, act[i] && (flg[i] & LOE_POS_END));
        int lower = 0;                                    // END positions in lower lanes
#pragma unroll
        for (int i = 0; i < SPL; ++i) { lower += __popc(m[i] & ((1u << lane) - 1)); n_end += __popc(m[i]); }
        int mine = 0;
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            if (act[i] && (flg[i] & LOE_POS_END)) { s_ends[lower + mine] = lane * SPL + i; ++mine; }
        }
    }
    __syncwarp();
    // lane w keeps END positions w, w+32, ... (n_end <= P <= 32*SPL)
    int my_end[SPL];
    bool end_ok[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const int w = lane + 32 * j;
        end_ok[j] = w < n_end;
        my_end[j] = end_ok[j] ? s_ends[w] : 0;
    }
    const bool loop = a.loop != 0;

    // gather d[] at this lane's END positions
    auto gather_ends = [&](float* out) {
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            const int owner = my_end[j] / SPL, slot = my_end[j] % SPL;
            float v = neg_inf();
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                const float t = __shfl_sync(FULL, d[i], owner);
                if (slot == i) v = t;
            }
            out[j] = end_ok[j] ? v : neg_inf();
        }
    };

    const float* __restrict__ sc = a.scores + f0 * a.ld;
    // t = 0
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        const float e0 = act[i] ? __ldg(sc + col[i]) : 0.f;
        d[i] = (act[i] && (flg[i] & LOE_POS_INIT)) ? __fadd_rn(e0, b0[i]) : neg_inf();
    }

    float ecur[kPre][SPL], enext[kPre][SPL];
#pragma unroll
    for (int k = 0; k < kPre; ++k)
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            const int t = 1 + k;
            ecur[k][i] = (act[i] && t < T) ? __ldg(sc + (int64_t)t * a.ld + col[i]) : 0.f;
        }
    uint32_t bits = 0;
    for (int tb = 1; tb < T; tb += kPre) {
#pragma unroll
        for (int k = 0; k < kPre; ++k)
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                const int t = tb + kPre + k;
                enext[k][i] = (act[i] && t < T) ? __ldg(sc + (int64_t)t * a.ld + col[i]) : 0.f;
            }
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const int t = tb + k;
            if (t < T) {
                // ---- cross-word candidate
                float cross32 = neg_inf(); double cross64 = -CUDART_INF; int cross_arg = 0;
                if (loop) {
                    float de[SPL];
                    gather_ends(de);
                    float lm = de[0];
#pragma unroll
                    for (int j = 1; j < SPL; ++j) lm = fmaxf(lm, de[j]);
                    const float m = fkey_inv(__reduce_max_sync(FULL, fkey(lm)));
                    int idx = -1;
                    if (a.pen_f64) {
                        cross64 = __dadd_rn(a.pen64, (double)m);
#pragma unroll
                        for (int j = 0; j < SPL; ++j) {
                            const unsigned eq = __ballot_sync(FULL, end_ok[j] && __dadd_rn(a.pen64, (double)de[j]) == cross64);
                            if (idx < 0 && eq) idx = 32 * j + __ffs(eq) - 1;
                        }
                    } else {
                        cross32 = __fadd_rn(a.pen32, m);
#pragma unroll
                        for (int j = 0; j < SPL; ++j) {
                            const unsigned eq = __ballot_sync(FULL, end_ok[j] && __fadd_rn(a.pen32, de[j]) == cross32);
                            if (idx < 0 && eq) idx = 32 * j + __ffs(eq) - 1;
                        }
                    }
                    if (idx < 0) idx = 0;
                    int arg_j = 0;
#pragma unroll
                    for (int j = 0; j < SPL; ++j) {
                        const int v = __shfl_sync(FULL, my_end[j], idx & 31);
                        if ((idx >> 5) == j) arg_j = v;
                    }
                    cross_arg = arg_j;
                    if (lane == 0) s_cross[t] = (uint8_t)cross_arg;
                }
                // ---- predecessors from the lane below
                float up1 = __shfl_up_sync(FULL, d[SPL - 1], 1);
                float up2 = (SPL >= 2) ? __shfl_up_sync(FULL, d[SPL >= 2 ? SPL - 2 : 0], 1) : __shfl_up_sync(FULL, d[0], 2);
                if (lane == 0) { up1 = neg_inf(); up2 = neg_inf(); }
                if (SPL == 1 && lane == 1) up2 = neg_inf();
                float nd[SPL];
#pragma unroll
                for (int i = 0; i < SPL; ++i) {
                    const float p1 = (i >= 1) ? d[i >= 1 ? i - 1 : 0] : up1;
                    const float p2 = (i >= 2) ? d[i >= 2 ? i - 2 : 0] : (i == 1 ? up1 : up2);
                    const float e = ecur[k][i];
                    float val; unsigned code;
                    if (loop && (flg[i] & LOE_POS_START)) {
                        const float selfc = __fadd_rn(b0[i], d[i]);
                        if (a.pen_f64) {
                            double mv = cross64; code = 3;
                            if ((double)selfc > mv) { mv = (double)selfc; code = 0; }
                            val = __double2float_rn(__dadd_rn(mv, (double)e));
                        } else {
                            float mv = cross32; code = 3;
                            if (selfc > mv) { mv = selfc; code = 0; }
                            val = __fadd_rn(mv, e);
                        }
                    } else {
                        float best = __fadd_rn(b2[i], p2); code = 2;
                        const float c1 = __fadd_rn(b1[i], p1);
                        if (c1 > best) { best = c1; code = 1; }
                        const float c0 = __fadd_rn(b0[i], d[i]);
                        if (c0 > best) { best = c0; code = 0; }
                        if (best == neg_inf()) code = 3;
                        val = __fadd_rn(best, e);
                    }
                    nd[i] = act[i] ? val : neg_inf();
                    bits |= code << (2 * ((t % FPW) * SPL + i));
                }
#pragma unroll
                for (int i = 0; i < SPL; ++i) d[i] = nd[i];
                if ((t % FPW) == FPW - 1 || t == T - 1) { s_bp[(t / FPW) * 32 + lane] = bits; bits = 0; }
            }
        }
#pragma unroll
        for (int k = 0; k < kPre; ++k)
#pragma unroll
            for (int i = 0; i < SPL; ++i) ecur[k][i] = enext[k][i];
    }
    __syncwarp();

    // ---- termination: best END (lowest index on ties), END scores out
    float de[SPL];
    gather_ends(de);
    if (a.end_scores) {
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
            const int w = lane + 32 * j;
            if (w < a.max_ends) a.end_scores[(int64_t)u * a.max_ends + w] = de[j];
        }
    }
    float lm = de[0];
#pragma unroll
    for (int j = 1; j < SPL; ++j) lm = fmaxf(lm, de[j]);
    const float m = fkey_inv(__reduce_max_sync(FULL, fkey(lm)));
    int bi = -1;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
        const unsigned eq = __ballot_sync(FULL, end_ok[j] && de[j] == m);
        if (bi < 0 && eq) bi = 32 * j + __ffs(eq) - 1;
    }
    if (bi < 0) bi = 0;
    if (lane == 0) { a.best[u] = bi; a.best_score[u] = (n_end > 0) ? m : neg_inf(); }

    // ---- backtrace (every lane walks the same chain; lane 0 records it)
    auto decode = [&](int t, int p) -> int {
        const uint32_t w = s_bp[(t / FPW) * 32 + p / SPL];
        const unsigned code = (w >> (2 * ((t % FPW) * SPL + p % SPL))) & 3u;
        if (code < 3) return p - (int)code;
        return (loop && (s_flags[p] & LOE_POS_START)) ? (int)s_cross[t] : 0;
    };
    if (T == 1) {
        if (lane == 0) s_path[0] = -1;
    } else {
        int prev = decode(T - 1, n_end > 0 ? s_ends[bi] : 0);
        if (lane == 0) s_path[T - 1] = (int8_t)prev;
        for (int t = T - 2; t >= 0; --t) {
            if (lane == 0) s_path[t] = (int8_t)prev;
            if (t >= 1) prev = decode(t, prev);
        }
    }
    __syncwarp();
    for (int t = lane; t < T; t += 32) a.path[f0 + t] = s_path[t];
}

template <int SPL>
static bool launch(const VitArgs& a, int n_utt, cudaStream_t s) {
    const WarpLayout L = warp_layout(a.max_frames, a.max_pos, SPL);
    const size_t smem = (size_t)L.total * kWarpsPerCta;
    if (smem > kWarpSmemCap) return false;
    static bool attr_done[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (dev >= 64 || !attr_done[dev]) {
        if (cudaFuncSetAttribute(viterbi_warp_kernel<SPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWarpSmemCap) != cudaSuccess)
            return false;
        if (dev < 64) attr_done[dev] = true;
    }
    viterbi_warp_kernel<SPL><<<(unsigned)((n_utt + kWarpsPerCta - 1) / kWarpsPerCta), kWarpsPerCta * 32, smem, s>>>(a, n_utt);
    return cudaGetLastError() == cudaSuccess;
}

bool viterbi_warp_launch(const VitArgs& a, int n_utt, cudaStream_t s) {
    if (a.max_pos <= 32) return launch<1>(a, n_utt, s);
    if (a.max_pos <= 64) return launch<2>(a, n_utt, s);
    return launch<4>(a, n_utt, s);
}

}  // namespace loe
