"""CPU: the index arithmetic the hand-written kernels rely on, restated in NumPy / checked on the source tables.

* the 10 x 16 decomposition of the 160-point complex FFT and the in-thread real-input post-pass of
  csrc/mfcc.cu (mfcc_mel_kernel): prime-factor 10-point DFT, twiddles, the rotated eleventh slot that lets the
  a = 0 thread pair bin k with 160 - k like every other thread;
* the real-input-first 20 x 16 decomposition of the 320-point real DFT of csrc/mfcc.cu (mfcc_mel_r_kernel): the
  prime-factor real 20-point DFT (five real 4-point DFTs, real 5-point DFTs for c = 0 and 2, one complex 5-point DFT
  for c = 1, conjugates for c = 3), the W_320 twiddles, one 16-point FFT per row, the mirrored bins, and the
  shared-memory addresses of the exchange rows and of the [plane][bin][frame] power spectra;
* the pairing table of csrc/emission_h16.cu (pair_chunk): the 8 MMAs of K = 16 cover the 15 split-operand chunk
  products exactly once, and no MMA is issued narrower than its chunks reach in the lower-triangular image;
* the fragment / accumulator / epilogue maps of csrc/kmeans.cu (accum2_kernel): every lane gathers exactly its own
  elements of the FP64 tensor-core fragments, and the packed statistics come out of the accumulator layout.
"""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "cs-304-speech-recognition-code_b200", "csrc")


def _w(n, e):
    return np.exp(-2j * np.pi * e / n)


def _dft5(v):
    c1, c2, s1, s2 = np.cos(2 * np.pi / 5), np.cos(4 * np.pi / 5), np.sin(2 * np.pi / 5), np.sin(4 * np.pi / 5)
    t1, t2, t3, t4 = v[1] + v[4], v[2] + v[3], v[1] - v[4], v[2] - v[3]
    m1, m2 = v[0] + c1 * t1 + c2 * t2, v[0] + c2 * t1 + c1 * t2
    q1, q2 = s1 * t3 + s2 * t4, s2 * t3 - s1 * t4
    return [v[0] + t1 + t2, m1 - 1j * q1, m2 - 1j * q2, m2 + 1j * q2, m1 + 1j * q1]


def _dft10(v):
    """Good-Thomas 2 x 5: input n = (5 na + 2 nb) mod 10, output k = (5 ka + 6 kb) mod 10 -- no twiddles."""
    c0 = _dft5([v[(2 * nb) % 10] for nb in range(5)])
    c1 = _dft5([v[(5 + 2 * nb) % 10] for nb in range(5)])
    out = [None] * 10
    for kb in range(5):
        out[(6 * kb) % 10] = c0[kb] + c1[kb]
        out[(5 + 6 * kb) % 10] = c0[kb] - c1[kb]
    return out


def _dft4(u):
    s0, s1, s2, s3 = u[0] + u[2], u[0] - u[2], u[1] + u[3], u[1] - u[3]
    return [s0 + s2, s1 - 1j * s3, s0 - s2, s1 + 1j * s3]


def _fft16(v):
    t = [[y * _w(16, b * c) for c, y in enumerate(_dft4([v[4 * a + b] for a in range(4)]))] for b in range(4)]
    out = [None] * 16
    for c in range(4):
        for d, y in enumerate(_dft4([t[b][c] for b in range(4)])):
            out[c + 4 * d] = y
    return out


def test_mel_kernel_fft_decomposition():
    rng = np.random.default_rng(0)
    x = rng.normal(0, 3000, 320)
    xw = x * (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(320) / 320))            # periodic Hann
    ref = np.abs(np.fft.rfft(xw)) ** 2
    z = xw[0::2] + 1j * xw[1::2]
    # step 1, thread n2: 10-point DFT over n1 of z[16 n1 + n2], twiddle W_160^(n2 k1); slot 10 = slot 0 * W_16^n2
    slots = np.zeros((11, 16), complex)
    for n2 in range(16):
        a = _dft10([z[16 * n1 + n2] for n1 in range(10)])
        assert np.allclose(a, np.fft.fft(z[n2::16]))
        for k1 in range(10):
            slots[k1, n2] = a[k1] * _w(160, n2 * k1)
        slots[10, n2] = slots[0, n2] * _w(16, n2)
    assert np.allclose(_fft16(np.arange(16) * (1 + 0.5j)), np.fft.fft(np.arange(16) * (1 + 0.5j)))
    # step 2, thread a: FFTs of slots a and 10 - a; bin k = a + 10 k2 pairs with 160 - k = (10 - a) + 10 (15 - k2)
    zfull = np.fft.fft(z)
    power = np.full(161, np.nan)
    for a in range(6):
        za, zb = _fft16(slots[a]), _fft16(slots[10 - a])
        for k2 in range(16):
            k = a + 10 * k2
            A, B = za[k2], zb[15 - k2]
            assert np.allclose(A, zfull[k % 160]) and np.allclose(B, zfull[(160 - k) % 160])
            e = complex(A.real + B.real, A.imag - B.imag)
            o = complex(A.imag + B.imag, B.real - A.real)
            t = _w(320, k) * o
            power[k], power[160 - k] = 0.25 * abs(e + t) ** 2, 0.25 * abs(e - t) ** 2
    assert not np.isnan(power).any()
    assert np.abs(power - ref).max() <= 1e-12 * ref.max()


def _rdft5(r):
    c1, c2, s1, s2 = np.cos(2 * np.pi / 5), np.cos(4 * np.pi / 5), np.sin(2 * np.pi / 5), np.sin(4 * np.pi / 5)
    t1, t2, t3, t4 = r[1] + r[4], r[2] + r[3], r[1] - r[4], r[2] - r[3]
    return r[0] + t1 + t2, r[0] + c1 * t1 + c2 * t2, s1 * t3 + s2 * t4, r[0] + c2 * t1 + c1 * t2, s2 * t3 - s1 * t4


def _rdft20(y):
    """mfcc_mel_r_kernel step 1: n1 = (5 a + 4 b) mod 20, k1 = c mod 4 = d mod 5; outputs k1 = 0..10."""
    u0, u1, u2 = [], [], []
    for b in range(5):
        y0, y1, y2, y3 = (y[(5 * a + 4 * b) % 20] for a in range(4))
        s0, s1, s2, s3 = y0 + y2, y0 - y2, y1 + y3, y1 - y3
        u0.append(s0 + s2); u2.append(s0 - s2); u1.append(complex(s1, -s3))
    Y = [None] * 11
    v0, m1, q1, m2, q2 = _rdft5(u0)
    Y[0], Y[4], Y[8] = complex(v0, 0), complex(m1, q1), complex(m2, q2)
    v0, m1, q1, m2, q2 = _rdft5(u2)
    Y[10], Y[6], Y[2] = complex(v0, 0), complex(m1, -q1), complex(m2, -q2)
    V = _dft5(u1)
    Y[5], Y[1], Y[9], Y[3], Y[7] = V[0], V[1], V[4], np.conj(V[2]), np.conj(V[3])
    return Y


def test_mel_r_kernel_real_first_decomposition():
    src = open(os.path.join(CSRC, "mfcc.cu")).read()
    const = {k: int(v) for k, v in re.findall(r"constexpr int (k\w+) = (\d+);", src[src.index("namespace r20 {"):])}
    frame_b, plane_b, batch = const["kFrameB"], const["kPlaneB"], const["kBatch"]
    row_b = batch * frame_b
    plane0 = 4 * row_b
    assert frame_b % 16 == 0 and (frame_b // 16) % 2 == 1          # 16-byte row loads of 8 frames: 8 different bank groups
    assert (plane_b // 4) % 32 == 16 and plane0 + 2 * plane_b <= 11 * row_b and 161 * 16 <= plane_b
    rng = np.random.default_rng(1)
    y = rng.normal(size=20)
    assert np.allclose(_rdft20(y), np.fft.fft(y)[:11])
    hann = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(320) / 320)
    area = {}
    frames = [rng.normal(0, 3000, 320) * hann for _ in range(batch)]
    # step 1, thread (frame, n2): rows k1 = 0..10 at byte k1 * row_b + fb * frame_b + n2 * 8
    rows = np.zeros((batch, 11, 16), complex)
    for fb, xw in enumerate(frames):
        for n2 in range(16):
            Y = _rdft20(xw[n2::16])
            for k1 in range(11):
                addr = k1 * row_b + fb * frame_b + n2 * 8
                assert addr not in area and addr + 8 <= 11 * row_b
                area[addr] = 1
                rows[fb, k1, n2] = Y[k1] * _w(320, n2 * k1)
    # step 2, thread (f8, k1): FFT16 of the row = bins k1 + 20 k2, mirrored above 160; float address in the planes
    for f8, xw in enumerate(frames):
        ref = np.abs(np.fft.rfft(xw)) ** 2
        power, written = np.full(161, np.nan), {}
        for k1 in range(11):
            X = _fft16(rows[f8, k1])
            for k2 in range(16):
                off = 4 * k1 + 80 * k2 if k2 < 8 else 4 * (320 - k1) - 80 * k2     # store_pw: lo[80 k2] / hi[-80 k2]
                b = off // 4
                assert b == (k1 + 20 * k2 if k1 + 20 * k2 <= 160 and k2 < 8 else 320 - k1 - 20 * k2) and 0 <= b <= 160
                byte = plane0 + (f8 >> 2) * plane_b + 4 * (off + (f8 & 3))
                assert plane0 <= byte < 11 * row_b
                power[b] = abs(X[k2]) ** 2
                written.setdefault(b, []).append(k1)
        assert not np.isnan(power).any() and np.abs(power - ref).max() <= 1e-12 * ref.max()
        assert all(len(set(v)) == 1 for v in written.values())                      # a bin is only ever rewritten by its own thread
    # store bank pattern: one instruction = fixed k2, lanes (f8 = 0..7, k1 = 4 pass + 0..3): 32 different banks
    for p_ in range(3):
        for k2 in range(16):
            banks = set()
            lanes = [(f8, 4 * p_ + kq) for kq in range(4) for f8 in range(8) if 4 * p_ + kq <= 10]
            for f8, k1 in lanes:
                off = 4 * k1 + 80 * k2 if k2 < 8 else 4 * (320 - k1) - 80 * k2
                banks.add(((plane0 + (f8 >> 2) * plane_b) // 4 + off + (f8 & 3)) % 32)
            assert len(banks) == len(lanes)


def test_mel_ex512_real_first_decomposition():
    """mel_ex512_kernel (csrc/mfcc_ex.cu): 512 = 32 x 16; the 32-point real DFT as a 16-point complex FFT of
    z[m] = (y[2m], y[2m+1]) plus the in-thread post-pass whose factor 1/2 rides in the W_512 twiddles; rows, mirrored
    bins and the shared-memory geometry."""
    src = open(os.path.join(CSRC, "mfcc_ex.cu")).read()
    const = {k: int(v) for k, v in re.findall(r"constexpr int (k\w+) = (\d+);", src[src.index("namespace r512 {"):])}
    frame_b, plane_b, batch, win_pitch = const["kFrameB"], const["kPlaneB"], const["kBatch"], const["kWinPitch"]
    row_b = batch * frame_b
    plane0 = 9 * row_b
    assert frame_b % 16 == 0 and (frame_b // 16) % 2 == 1
    assert (plane_b // 4) % 32 == 16 and plane0 + 2 * plane_b <= 17 * row_b and 257 * 16 <= plane_b
    assert win_pitch % 2 == 0 and (win_pitch // 2) % 2 == 1 and win_pitch >= 32     # 8-byte loads of 16 lanes: 32 banks
    rng = np.random.default_rng(2)

    def rdft32_doubled(y):
        """what the thread holds before the W_512 twiddles: Y[0], Y[16] plain, 2 Y[k] for k = 1..15"""
        Z = np.array(_fft16(list(y[0::2] + 1j * y[1::2])))
        Y = np.zeros(17, complex)
        Y[0], Y[16], Y[8] = Z[0].real + Z[0].imag, Z[0].real - Z[0].imag, 2 * np.conj(Z[8])
        for k in range(1, 8):
            A, B = Z[k], Z[16 - k]
            e, o = complex(A.real + B.real, A.imag - B.imag), complex(A.imag + B.imag, B.real - A.real)
            wo = _w(32, k) * o
            Y[k], Y[16 - k] = e + wo, np.conj(e - wo)
        return Y
    y = rng.normal(size=32)
    ref32 = np.fft.fft(y)[:17]
    got = rdft32_doubled(y)
    assert np.allclose(got[[0, 16]], ref32[[0, 16]]) and np.allclose(got[1:16], 2 * ref32[1:16])
    xw = rng.normal(0, 3000, 512) * np.hamming(512)
    rows = np.zeros((17, 16), complex)
    for n2 in range(16):
        Y = rdft32_doubled(xw[n2::16])
        for k1 in range(17):
            rows[k1, n2] = Y[k1] * _w(512, n2 * k1) * (0.5 if 1 <= k1 <= 15 else 1.0)
    ref = np.abs(np.fft.rfft(xw)) ** 2
    power = np.full(257, np.nan)
    for k1 in range(17):
        X = _fft16(rows[k1])
        for k2 in range(16):
            off = 4 * k1 + 128 * k2 if k2 < 8 else 4 * (512 - k1) - 128 * k2        # store_pw: lo[128 k2] / hi[-128 k2]
            b = off // 4
            assert b == (k1 + 32 * k2 if k2 < 8 else 512 - k1 - 32 * k2) and 0 <= b <= 256
            power[b] = abs(X[k2]) ** 2
    assert not np.isnan(power).any() and np.abs(power - ref).max() <= 1e-12 * ref.max()
    # step 2 passes: rows 13..16 (kept in registers), 9..12, 5..8, 1..4, 0 -- the planes only cover rows 9..16
    assert plane0 == 9 * row_b and [13 + kq for kq in range(4)][-1] == 16


def test_emission_h16_pairing_table():
    src = open(os.path.join(CSRC, "emission_h16.cu")).read()
    body = src[src.index("constexpr int t[kNumMma][5] = {"):]
    body = body[:body.index("};")]
    rows = [tuple(int(v) for v in m) for m in re.findall(r"\{(\d+), (\d+), (\d+), (\d+), (\d+)\}", body)]
    assert len(rows) == 8

    def a_part(c):      # A chunks: hi 0-4, lo 5-9, zero 10
        return ("hi", c) if c < 5 else ("lo", c - 5) if c < 10 else ("zero", None)

    def b_part(c):      # B chunks: hi 0-4, lo 5-9, second copy of hi 10-14
        return ("hi", c) if c < 5 else ("lo", c - 5) if c < 10 else ("hi", c - 10)

    products = []
    for a0, a1, b0, b1, blocks in rows:
        assert 1 <= blocks <= 5
        for ac, bc in ((a0, b0), (a1, b1)):
            (pa, ca), (pb, cb) = a_part(ac), b_part(bc)
            if pa == "zero":
                continue                                    # zero x anything finite
            assert ca == cb                                 # same K chunk on both sides
            assert blocks >= ca + 1                         # the MMA covers every column that chunk reaches (j <= 8 c + 7)
            products.append((pa, pb, ca))
    want = [(pa, pb, c) for c in range(5) for pa, pb in (("hi", "hi"), ("lo", "hi"), ("hi", "lo"))]
    assert sorted(products) == sorted(want)                 # hi*hi + lo*hi + hi*lo of every chunk, once each
    assert rows[0][4] == 5                                  # the first MMA (accumulate = 0) overwrites all 240 columns
    assert sum(r[4] for r in rows) * 48 == 1152             # 60 % of the dense 8 x 240 (products paired by width)


def _stockham_real_power(x, N):
    """NumPy walk through mel_ex_kernel's FFT (csrc/mfcc_ex.cu): z[n] = x[2n] + i x[2n+1], Stockham autosort passes of
    radix 4 (sub-transform size Ns = 1, 4, ...) plus one radix-2 pass when log2(N/2) is odd, with the kernel's index
    arithmetic (j0 = ((j - k) << 2) + k, twiddle index r k M / (4 Ns)), then the real-input post-pass."""
    M = N // 2
    wm = np.exp(-2j * np.pi * np.arange(M) / M)
    wn = np.exp(-2j * np.pi * np.arange(M + 1) / N)
    a = x[0::2] + 1j * x[1::2]
    Ns = 1
    while Ns * 4 <= M:
        out = np.zeros(M, dtype=complex)
        step = M // (Ns * 4)
        for j in range(M // 4):
            k = j & (Ns - 1)
            v0, v1, v2, v3 = a[j], a[j + M // 4], a[j + M // 2], a[j + 3 * M // 4]
            if Ns > 1:
                v1 *= wm[k * step]; v2 *= wm[2 * k * step]; v3 *= wm[3 * k * step]
            s0, s1, s2, s3 = v0 + v2, v0 - v2, v1 + v3, v1 - v3
            j0 = ((j - k) << 2) + k
            out[j0] = s0 + s2
            out[j0 + Ns] = s1 - 1j * s3
            out[j0 + 2 * Ns] = s0 - s2
            out[j0 + 3 * Ns] = s1 + 1j * s3
        a = out
        Ns *= 4
    if Ns < M:
        out = np.zeros(M, dtype=complex)
        for j in range(M // 2):
            v0, v1 = a[j], a[j + M // 2] * wm[j]
            out[j] = v0 + v1
            out[j + Ns] = v0 - v1
        a = out
    pw = np.zeros(M + 1)
    for k in range(M + 1):
        A, B = a[k & (M - 1)], a[(M - k) & (M - 1)]
        e = complex(A.real + B.real, A.imag - B.imag)
        o = complex(A.imag + B.imag, B.real - A.real)
        xk = e + wn[k] * o
        pw[k] = 0.25 * abs(xk) ** 2
    return pw


def test_parameterised_fft_index_arithmetic():
    rng = np.random.default_rng(0)
    for N in (64, 128, 256, 512, 1024):
        x = rng.normal(size=N)
        ref = np.abs(np.fft.rfft(x)) ** 2
        got = _stockham_real_power(x, N)
        assert np.allclose(got, ref, rtol=1e-10, atol=1e-10), N


def test_reversed_cholesky_gives_the_lower_triangular_whitening_matrix():
    """csrc/mstep.cu: M = chol(J cov J) (J = exchange matrix), W[k][j] = M^-1[38 - j][38 - k].  W must be LOWER triangular
    (what the 3xFP16 emission image needs) with W W^T = cov^-1, and log|cov| = 2 sum log diag(M)."""
    rng = np.random.default_rng(1)
    D = 39
    q, _ = np.linalg.qr(rng.normal(size=(D, D)))
    cov = (q * np.logspace(-3, 2, D)) @ q.T
    cov = (cov + cov.T) / 2
    J = np.eye(D)[::-1]
    M = np.linalg.cholesky(J @ cov @ J)
    Minv = np.linalg.inv(M)
    W = np.array([[Minv[D - 1 - j, D - 1 - k] for j in range(D)] for k in range(D)])
    assert np.allclose(np.triu(W, 1), 0.0)
    assert np.allclose(W @ W.T, np.linalg.inv(cov), rtol=1e-8, atol=1e-8 * np.abs(np.linalg.inv(cov)).max())
    assert np.isclose(2 * np.log(np.diag(M)).sum(), np.linalg.slogdet(cov)[1])
    x = rng.normal(size=(5, D)); mu = rng.normal(size=D)
    ref = np.einsum("ti,ij,tj->t", x - mu, np.linalg.inv(cov), x - mu)
    assert np.allclose(np.sum(((x - mu) @ W) ** 2, axis=1), ref, rtol=1e-9)


def test_accum2_fragments_come_straight_from_the_gather():
    """csrc/kmeans.cu (accum2_kernel): with Y the 4 x 40 block of a step (rows [x - shift, 1]), lane l feeds
    Y[l & 3][8 b + (l >> 2)] to the DMMAs as its element of BOTH the A fragment (A[m][k] = Y[k][8 bi + m], lane holds
    A[l >> 2][l & 3]) and the B fragment (B[k][n] = Y[k][8 bj + n], lane holds B[l & 3][l >> 2]) of tile (bi, bj) of
    Y^T Y; the accumulator layout D[l >> 2][2 (l & 3) + h] and the (tile, half, lane) -> (i, j) map of the epilogue
    must then reproduce the packed statistics [N | sum y | upper triangle of sum y y^T]."""
    rng = np.random.default_rng(3)
    D, n_frames = 39, 10                                    # 10 frames: two full steps and a partial one
    x = rng.normal(size=(n_frames, D))
    shift = rng.normal(size=D)
    tiles = [(bi, bj) for bi in range(5) for bj in range(bi, 5)]
    c = np.zeros((15, 2, 32))
    for step in range((n_frames + 3) // 4):
        f = np.zeros((32, 5))                               # f[lane][b]
        for lane in range(32):
            row, col0 = lane & 3, lane >> 2
            fr = 4 * step + row
            for b in range(5):
                k = col0 + 8 * b
                pre = (x[fr, k] if k < D else 1.0) if fr < n_frames else (1.0 if b == 4 else 0.0)     # what gather() leaves
                v = pre - (shift[k] if k < D else 0.0)
                f[lane, b] = v if fr < n_frames else 0.0                                               # the partial-step select
        for t, (bi, bj) in enumerate(tiles):
            A = np.zeros((8, 4)); B = np.zeros((4, 8))
            for lane in range(32):
                A[lane >> 2, lane & 3] = f[lane, bi]
                B[lane & 3, lane >> 2] = f[lane, bj]
            Dm = A @ B
            for lane in range(32):
                for h in range(2):
                    c[t, h, lane] += Dm[lane >> 2, 2 * (lane & 3) + h]
    y = np.concatenate((x - shift, np.ones((n_frames, 1))), axis=1)
    full = y.T @ y
    stride = 1 + D + D * (D + 1) // 2
    out = np.zeros(stride)
    for e in range(stride):
        i = j = D
        if 1 <= e <= D:
            i, j = e - 1, D
        elif e > D:
            r, row = e - 1 - D, 0
            while r >= D - row:
                r -= D - row
                row += 1
            i, j = row, row + r
        bi, bj = i >> 3, j >> 3
        t = bi * 5 - (bi * (bi - 1)) // 2 + (bj - bi)
        assert tiles[t] == (bi, bj)
        out[e] = c[t, j & 1, ((i & 7) << 2) + ((j & 7) >> 1)]
    iu = np.triu_indices(D)
    want = np.concatenate(([n_frames], (x - shift).sum(0), ((x - shift).T @ (x - shift))[iu]))
    assert np.allclose(out, want, rtol=1e-12, atol=1e-12)
    assert np.allclose(full[D, D], n_frames)
