// Split of one feature row into the binary16 hi / lo A operand of the 3xFP16 emission kernel (emission_h16.cu), shared by
// that kernel's producer warps and by the cepstrum kernel (mfcc.cu), which can write the same operand straight into a
// global image so that the emission kernel only has to bulk-copy it (no producer work, no second pass over the features).
#pragma once
#include <cuda_fp16.h>
#include <math_constants.h>
#include <stdint.h>

namespace loe {
namespace h16 {

constexpr int kStageK = 40;                 // 39 features + the constant 1 of the bias row
constexpr int kStageChunksPerPart = 5;      // 16-byte chunks (8 halfs) of the hi part, and of the lo part
constexpr int kStageLbo = 128 * 16;         // bytes between K-adjacent chunks of a 128-row tile
constexpr int kImgTileBytes = 2 * kStageChunksPerPart * kStageLbo;     // 20 480: hi chunks 0-4, lo chunks 5-9

// hi/lo split of 8 consecutive values into one 16-byte chunk each (packed conversions: F2FP converts two
// values per instruction, the scalar F2F runs on the slow conversion pipe)
__device__ __forceinline__ void split_store(const float* x, uint8_t* a_row, int kc) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __half2 hh = __floats2half2_rn(x[2 * q], x[2 * q + 1]);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(x[2 * q] - hf.x, x[2 * q + 1] - hf.y);
        h[q] = *reinterpret_cast<const uint32_t*>(&hh);
        l[q] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(a_row + kc * kStageLbo) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(a_row + (kStageChunksPerPart + kc) * kStageLbo) = make_uint4(l[0], l[1], l[2], l[3]);
}

// One thread stages one feature row (39 values + the constant 1 of the bias row) at a_row (= tile base + row * 16).
// Returns 4^e of the power-of-two scale 2^-e applied to the row: 1 unless its largest magnitude reaches 2^15.
__device__ __forceinline__ float stage_row(const float* __restrict__ row, uint8_t* __restrict__ a_row) {
    float v[kStageK];
    float mx = 0.0f;
#pragma unroll
    for (int c = 0; c < kStageK - 1; ++c) {
        v[c] = row[c];
        mx = fmaxf(mx, fabsf(v[c]));
    }
    v[kStageK - 1] = 1.0f;
    float inv2 = 1.0f;
    if (!(mx < 32768.0f)) {                        // rare; also taken for NaN
        const int e = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 127 - 14;       // 1 .. 114
        const float scale = __uint_as_float((uint32_t)(127 - e) << 23);
        inv2 = (2 * e < 128) ? __uint_as_float((uint32_t)(127 + 2 * e) << 23) : CUDART_INF_F;
#pragma unroll
        for (int c = 0; c < kStageK; ++c) v[c] *= scale;
    }
#pragma unroll
    for (int kc = 0; kc < kStageChunksPerPart; ++kc) split_store(v + kc * 8, a_row, kc);
    return inv2;
}

}  // namespace h16
}  // namespace loe
