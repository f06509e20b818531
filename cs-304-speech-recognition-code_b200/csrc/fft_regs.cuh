// Register-resident FFT building blocks shared by the MFCC kernels (mfcc.cu, mfcc_ex.cu): complex arithmetic on packed
// f32x2 register pairs, 4- / 5- / 10- / 16-point transforms held entirely in the registers of one thread, and the exact
// int16 -> float32 conversion.
#pragma once
#include "common.cuh"

namespace loe {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// complex add / subtract / element-wise multiply as ONE packed instruction (add / sub / mul .f32x2: both halves
// IEEE round-to-nearest, i.e. the same results as two scalar operations, half the issue slots)
__device__ __forceinline__ unsigned long long pack2(float2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(a)), "l"(pack2(b)));
    return unpack2(r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(a)), "l"(pack2(b)));
    return unpack2(r);
}
__device__ __forceinline__ float2 emul(float2 a, float2 b) {            // (a.x * b.x, a.y * b.y), never contracted
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(a)), "l"(pack2(b)));
    return unpack2(r);
}

// forward 5-point DFT
__device__ __forceinline__ void dft5(float2 v0, float2 v1, float2 v2, float2 v3, float2 v4, float2* y) {
    const float C1 = 0.30901699437494745f, C2 = -0.80901699437494745f;   // cos(2pi/5), cos(4pi/5)
    const float S1 = 0.95105651629515353f, S2 = 0.58778525229247314f;    // sin(2pi/5), sin(4pi/5)
    const float2 t1 = cadd(v1, v4), t2 = cadd(v2, v3), t3 = csub(v1, v4), t4 = csub(v2, v3);
    y[0] = cadd(cadd(v0, t1), t2);
    const float2 m1 = make_float2(v0.x + C1 * t1.x + C2 * t2.x, v0.y + C1 * t1.y + C2 * t2.y);
    const float2 m2 = make_float2(v0.x + C2 * t1.x + C1 * t2.x, v0.y + C2 * t1.y + C1 * t2.y);
    const float2 q1 = make_float2(S1 * t3.x + S2 * t4.x, S1 * t3.y + S2 * t4.y);
    const float2 q2 = make_float2(S2 * t3.x - S1 * t4.x, S2 * t3.y - S1 * t4.y);
    y[1] = make_float2(m1.x + q1.y, m1.y - q1.x);
    y[4] = make_float2(m1.x - q1.y, m1.y + q1.x);
    y[2] = make_float2(m2.x + q2.y, m2.y - q2.x);
    y[3] = make_float2(m2.x - q2.y, m2.y + q2.x);
}

// forward 10-point DFT, prime-factor (Good-Thomas) 2 x 5: no twiddles between the two stages.
//   input n = (5 na + 2 nb) mod 10, output k = (5 ka + 6 kb) mod 10
__device__ __forceinline__ void dft10(const float2* v, float2* out) {
    float2 c0[5], c1[5];
    dft5(v[0], v[2], v[4], v[6], v[8], c0);
    dft5(v[5], v[7], v[9], v[1], v[3], c1);
    out[0] = cadd(c0[0], c1[0]); out[5] = csub(c0[0], c1[0]);
    out[6] = cadd(c0[1], c1[1]); out[1] = csub(c0[1], c1[1]);
    out[2] = cadd(c0[2], c1[2]); out[7] = csub(c0[2], c1[2]);
    out[8] = cadd(c0[3], c1[3]); out[3] = csub(c0[3], c1[3]);
    out[4] = cadd(c0[4], c1[4]); out[9] = csub(c0[4], c1[4]);
}

__device__ __forceinline__ void dft4(float2 u0, float2 u1, float2 u2, float2 u3, float2& y0, float2& y1, float2& y2, float2& y3) {
    const float2 s0 = cadd(u0, u2), s1 = csub(u0, u2), s2 = cadd(u1, u3), s3 = csub(u1, u3);
    y0 = cadd(s0, s2);
    y1 = make_float2(s1.x + s3.y, s1.y - s3.x);     // s1 - i s3
    y2 = csub(s0, s2);
    y3 = make_float2(s1.x - s3.y, s1.y + s3.x);     // s1 + i s3
}

// forward 16-point FFT in place, natural order in and out: n = 4a + b, k = c + 4d
__device__ __forceinline__ void fft16(float2* v) {
    const float CA = 0.92387953251128674f, SA = 0.38268343236508977f, R = 0.70710678118654752f;
    float2 t[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4(v[b], v[4 + b], v[8 + b], v[12 + b], t[b][0], t[b][1], t[b][2], t[b][3]);
    // twiddles W_16^(b c)
    t[1][1] = cmul(t[1][1], make_float2(CA, -SA));                       // W^1
    t[1][2] = make_float2(R * (t[1][2].x + t[1][2].y), R * (t[1][2].y - t[1][2].x));   // W^2 = (1 - i) / sqrt 2
    t[1][3] = cmul(t[1][3], make_float2(SA, -CA));                       // W^3
    t[2][1] = make_float2(R * (t[2][1].x + t[2][1].y), R * (t[2][1].y - t[2][1].x));   // W^2
    t[2][2] = make_float2(t[2][2].y, -t[2][2].x);                        // W^4 = -i
    t[2][3] = make_float2(R * (t[2][3].y - t[2][3].x), -R * (t[2][3].x + t[2][3].y));  // W^6 = (-1 - i) / sqrt 2
    t[3][1] = cmul(t[3][1], make_float2(SA, -CA));                       // W^3
    t[3][2] = make_float2(R * (t[3][2].y - t[3][2].x), -R * (t[3][2].x + t[3][2].y));  // W^6
    t[3][3] = cmul(t[3][3], make_float2(-CA, SA));                       // W^9
#pragma unroll
    for (int c = 0; c < 4; ++c) dft4(t[0][c], t[1][c], t[2][c], t[3][c], v[c], v[c + 4], v[c + 8], v[c + 12]);
}

// sample -> float.  int16 goes through the 1.5 * 2^23 magic number (integer add + float subtract on the main pipes,
// exact for |s| < 2^22) instead of I2F, which issues at a quarter of the rate
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(short v) { return __int_as_float(0x4B400000 + (int)v) - 12582912.0f; }

}  // namespace loe
