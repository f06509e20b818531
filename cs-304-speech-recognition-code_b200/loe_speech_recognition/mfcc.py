"""MFCC front end -- same public API as the reference's ``mfcc.py`` (MFCC dataclass,
``feature_vector``, ``normalize_mfccs``, ``batch``), computed by the sm_100a kernels in
csrc/mfcc.cu through ``loe_mfcc_dev`` (include/loe_b200.h).

Reference: src/loe_speech_recognition/mfcc.py:12-84.  The reference delegates the arithmetic
to librosa (melspectrogram n_mels=40 n_fft=320 hop=160 fmin=133.33 fmax=6855.4976,
power_to_db(ref=np.max), mfcc(n_mfcc=13), delta, delta order 2); the kernels restate that
pipeline.  Only the filterbank weights are prepared on the host (once per
sample rate).
"""
from __future__ import annotations

import functools

from dataclasses import dataclass, field
from typing import List

import numpy as np
from numpy.typing import NDArray

from . import _native

N_FFT = 320
HOP = 160
N_MELS = 40
FMIN = 133.33
FMAX = 6855.4976


def _slaney_hz_to_mel(f: float) -> float:
    f_sp = 200.0 / 3
    if f >= 1000.0:
        return 1000.0 / f_sp + np.log(f / 1000.0) / (np.log(6.4) / 27.0)
    return f / f_sp


def _slaney_mel_to_hz(m: np.ndarray) -> np.ndarray:
    f_sp = 200.0 / 3
    lin = f_sp * m
    log = 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 1000.0 / f_sp))
    return np.where(m >= 1000.0 / f_sp, log, lin)


def mel_filterbank(sample_rate: float) -> NDArray[np.float32]:
    """Dense (40, 161) slaney-normalised triangular filterbank (librosa.filters.mel semantics)."""
    n_bins = 1 + N_FFT // 2
    freqs = np.arange(n_bins, dtype=np.float64) * (float(sample_rate) / N_FFT)
    edges = _slaney_mel_to_hz(np.linspace(_slaney_hz_to_mel(FMIN), _slaney_hz_to_mel(FMAX), N_MELS + 2))
    width = np.diff(edges)
    w = np.zeros((N_MELS, n_bins), dtype=np.float32)
    for m in range(N_MELS):
        rise = (freqs - edges[m]) / width[m]
        fall = (edges[m + 2] - freqs) / width[m + 1]
        w[m] = np.maximum(0.0, np.minimum(rise, fall))
    w *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return w


def _distinct_bank_starts(first, last, span, lanes_per_filter):
    """Window starts w[f] <= first[f] with w[f] + span > last[f] such that the addresses the 32 lanes read in one
    iteration -- w[f] + j, j < lanes_per_filter -- fall into 32 different shared-memory banks (residues mod 32),
    or None.  Depth-first search with the natural start tried first; 32 / lanes_per_filter filters."""
    n = len(first)
    cands = [[w for w in range(first[f], last[f] - span, -1) if w >= 0] for f in range(n)]
    order = sorted(range(n), key=lambda f: len(cands[f]))          # most constrained filters first
    chosen = [None] * n
    budget = [200000]

    def place(i, used):
        if i == n:
            return True
        f = order[i]
        for w in cands[f]:
            budget[0] -= 1
            if budget[0] < 0:
                return False
            rs = {(w + j) % 32 for j in range(lanes_per_filter)}
            if rs & used:
                continue
            chosen[f] = w
            if place(i + 1, used | rs):
                return True
        return False

    return chosen if place(0, set()) else None


def _quarter_wavefronts(starts, first, last, span, lanes_per_filter, predicated=False):
    """Shared-memory wavefronts of one quarter-warp walking its windows in the [bin][4 frames] power layout of
    mfcc_mel_r_kernel (16 bytes per bin): per iteration the largest number of DIFFERENT bins that share a 16-byte bank
    group (bin mod 8).  The kernel loads every entry of a window (zero weights included); ``predicated`` counts only
    the entries inside the filters' supports."""
    total = 0
    for it in range(span // lanes_per_filter):
        groups = {}
        for f, w in enumerate(starts):
            for j in range(lanes_per_filter):
                b = w + j + lanes_per_filter * it
                if not predicated or first[f] <= b <= last[f]:
                    groups.setdefault(b % 8, set()).add(b)
        total += max((len(v) for v in groups.values()), default=0)
    return total


def _min_wavefront_starts(first, last, span, lanes_per_filter, restarts=40):
    """Window starts (w[f] <= first[f], w[f] + span > last[f], w[f] >= 0) for the 16-byte-per-bin layout: per
    quarter-warp (8 lanes = 8 / lanes_per_filter filters) coordinate descent on _quarter_wavefronts from the natural
    starts and from seeded random starts (deterministic): at 16 kHz 69 wavefronts per 8 frames and plane where the
    natural starts need 139 (64 = one per iteration and quarter-warp is the floor)."""
    import random
    rnd = random.Random(304)
    n = len(first)
    per_q = 8 // lanes_per_filter
    chosen = list(first)
    for q0 in range(0, n, per_q):
        fs = list(range(q0, min(n, q0 + per_q)))
        cands = [[w for w in range(first[f], last[f] - span, -1) if w >= 0] for f in fs]
        fi, la = [first[f] for f in fs], [last[f] for f in fs]
        best = None
        for r in range(restarts):
            cur = list(fi) if r == 0 else [rnd.choice(c) for c in cands]
            cost = _quarter_wavefronts(cur, fi, la, span, lanes_per_filter)
            improved = True
            while improved:
                improved = False
                for i in range(len(fs)):
                    for w in cands[i]:
                        trial = cur[:i] + [w] + cur[i + 1:]
                        c = _quarter_wavefronts(trial, fi, la, span, lanes_per_filter)
                        if c < cost:
                            cost, cur, improved = c, trial, True
            if best is None or cost < best[0]:
                best = (cost, cur)
        for f, w in zip(fs, best[1]):
            chosen[f] = w
    return chosen


@functools.lru_cache(maxsize=8)
def mel_lane_tables(sample_rate: float):
    """The filterbank in the lane-balanced layout loe_mfcc_dev reads (include/loe_b200.h):
    (bin int32 [(na+nb)*32], weight float32 [(na+nb)*32], na, nb).  Round A: lane l owns filter l
    (filters 0..31), one bin per iteration.  Round B: lanes 4q..4q+3 share filter 32+q (the
    widest filters), lane 4q + j takes every fourth bin starting at its window start + j.

    Every lane walks CONSECUTIVE bins (round A: start + it; round B: start + j + 4*it), so the kernel only
    needs the first bin of each lane; entries outside the filter's support carry weight 0.  A filter narrower
    than the na (4 nb) bins of its window leaves room to slide the window: the starts are chosen such that
    the 32 addresses of one iteration hit 32 different shared-memory banks (the natural starts, 3, 4, 6, ...,
    69 at 16 kHz, collide two-way in every iteration)."""
    dense = mel_filterbank(sample_rate)
    nz = [np.nonzero(dense[m])[0] for m in range(N_MELS)]
    na = max((len(nz[m]) for m in range(32)), default=0)
    nb = max(((len(nz[m]) + 3) // 4 for m in range(32, N_MELS)), default=0)
    if na > _native.LOE_MEL_NA_MAX or nb > _native.LOE_MEL_NB_MAX:
        raise NotImplementedError(f"mel filters too wide for the kernel tables at sample_rate={sample_rate}")
    first = [int(nz[m][0]) if len(nz[m]) else 0 for m in range(N_MELS)]
    last = [int(nz[m][-1]) if len(nz[m]) else 0 for m in range(N_MELS)]
    for m in range(N_MELS):
        assert len(nz[m]) == 0 or np.array_equal(nz[m], np.arange(first[m], last[m] + 1))
    if (na, nb) == (11, 5):
        # the 16 kHz table: mfcc_mel_r_kernel reads 16-byte [bin][4 frames] entries, a quarter-warp per wavefront
        start_a = _min_wavefront_starts(first[:32], last[:32], na, 1)
        start_b = _min_wavefront_starts(first[32:], last[32:], 4 * nb, 4)
    else:
        start_a = _distinct_bank_starts(first[:32], last[:32], na, 1) or first[:32]
        start_b = _distinct_bank_starts(first[32:], last[32:], 4 * nb, 4) or first[32:]
    bins = np.zeros((na + nb, 32), dtype=np.int32)
    w = np.zeros((na + nb, 32), dtype=np.float32)
    n_bins = dense.shape[1]
    for m in range(32):
        for j in range(na):
            k = start_a[m] + j
            bins[j, m] = k
            if k < n_bins:
                w[j, m] = dense[m, k]
    for q, m in enumerate(range(32, N_MELS)):
        for j in range(4 * nb):
            k = start_b[q] + j
            bins[na + j // 4, 4 * q + j % 4] = k
            if k < n_bins:
                w[na + j // 4, 4 * q + j % 4] = dense[m, k]
    # nothing of a filter's support may fall outside its window
    assert np.isclose(w[:na].sum(), dense[:32].sum(), rtol=1e-6) and np.isclose(w[na:].sum(), dense[32:].sum(), rtol=1e-6)
    return np.ascontiguousarray(bins.reshape(-1)), np.ascontiguousarray(w.reshape(-1)), int(na), int(nb)


@dataclass(frozen=True)
class MFCCConfig:
    """Front-end parameters (added; the reference hard-wires REFERENCE, mfcc.py:31-34).  ``MFCC.batch(..., config=)``
    runs any other set through ``loe_mfcc_ex_dev`` (csrc/mfcc_ex.cu): power-of-two FFT sizes up to 1024, a window
    shorter than the FFT (centred, zero-padded, as librosa pads it), pre-emphasis, dB or natural log, and the
    normalisation of the static block.  ``MFCCConfig.spec()`` is BASELINE.json configs[3]."""
    n_fft: int = N_FFT
    win_length: int = N_FFT
    hop_length: int = HOP
    window: str = "hann"                 # "hann" | "hamming" (periodic, scipy.signal.get_window(fftbins=True))
    n_mels: int = N_MELS
    fmin: float = FMIN
    fmax: float = FMAX
    preemphasis: float = 0.0
    log: str = "db"                      # "db": power_to_db(ref=np.max), top_db 80 | "ln": natural log
    n_mfcc: int = 13
    norm: str = "frame"                  # "frame" (mfcc.py:62-66) | "cmn" | "cmvn" | "none"

    @classmethod
    def spec(cls) -> "MFCCConfig":
        """25 ms / 10 ms frames at 16 kHz, 512-point FFT, Hamming, pre-emphasis 0.97, 40 mel, 13 ceps + deltas, CMN."""
        return cls(n_fft=512, win_length=400, window="hamming", preemphasis=0.97, log="ln", norm="cmn")

    @property
    def is_reference(self) -> bool:
        return self == MFCCConfig()

    def window_table(self) -> NDArray[np.float32]:
        n = np.arange(self.win_length, dtype=np.float64)
        if self.window == "hann":
            w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / self.win_length)
        elif self.window == "hamming":
            w = 0.54 - 0.46 * np.cos(2.0 * np.pi * n / self.win_length)
        else:
            raise ValueError(f"unknown window {self.window!r}")
        if self.win_length > self.n_fft:
            raise ValueError("win_length must not exceed n_fft")
        lpad = (self.n_fft - self.win_length) // 2
        out = np.zeros(self.n_fft, dtype=np.float32)
        out[lpad:lpad + self.win_length] = w
        return out

    def mel_tables(self, sample_rate: float):
        """(start int32 [n_mels], length int32 [n_mels], weights float32 [n_mels, pitch]) of the slaney filterbank."""
        n_bins = 1 + self.n_fft // 2
        freqs = np.arange(n_bins, dtype=np.float64) * (float(sample_rate) / self.n_fft)
        edges = _slaney_mel_to_hz(np.linspace(_slaney_hz_to_mel(self.fmin), _slaney_hz_to_mel(self.fmax), self.n_mels + 2))
        width = np.diff(edges)
        dense = np.zeros((self.n_mels, n_bins), dtype=np.float32)
        for m in range(self.n_mels):
            rise = (freqs - edges[m]) / width[m]
            fall = (edges[m + 2] - freqs) / width[m + 1]
            dense[m] = np.maximum(0.0, np.minimum(rise, fall))
        dense *= (2.0 / (edges[2:] - edges[:-2]))[:, None].astype(np.float32)
        start = np.zeros(self.n_mels, dtype=np.int32)
        length = np.zeros(self.n_mels, dtype=np.int32)
        for m in range(self.n_mels):
            nz = np.nonzero(dense[m])[0]
            if len(nz):
                start[m], length[m] = nz[0], nz[-1] - nz[0] + 1
        first = start.copy()
        if self.n_fft == 512 and self.n_mels > 0 and int(length.min()) > 0:
            # mel_ex512_kernel reads 16-byte [bin][4 frames] entries, a quarter-warp per wavefront, and every lane walks a
            # window as long as the widest filter of its round: a narrower filter may start earlier (leading zero weights)
            # -- the starts are slid so that the loads of one iteration spread over the bank groups (same search as
            # mel_lane_tables).  Round A: filters 0..31, one lane each; round B: the others, lb lanes each.
            n_a = min(self.n_mels, 32)
            n_b = self.n_mels - n_a
            last = start + length - 1
            span_a = int(length[:n_a].max())
            start[:n_a] = _min_wavefront_starts([int(v) for v in first[:n_a]], [int(v) for v in last[:n_a]], span_a, 1)
            if n_b:
                lb = 32
                while lb > 1 and lb * n_b > 32:
                    lb >>= 1
                if lb <= 8:
                    span_b = lb * int(-(-int(length[n_a:].max()) // lb))
                    start[n_a:] = _min_wavefront_starts([int(v) for v in first[n_a:]], [int(v) for v in last[n_a:]], span_b, lb)
            length = (last - start + 1).astype(np.int32)
        pitch = max(1, int(length.max()))
        w = np.zeros((self.n_mels, pitch), dtype=np.float32)
        for m in range(self.n_mels):
            w[m, :length[m]] = dense[m, start[m]:start[m] + length[m]]
        return start, length, w

    def dct_table(self) -> NDArray[np.float32]:
        k = np.arange(self.n_mfcc, dtype=np.float64)[:, None]
        n = np.arange(self.n_mels, dtype=np.float64)[None, :]
        d = np.sqrt(2.0 / self.n_mels) * np.cos(np.pi * (2 * n + 1) * k / (2.0 * self.n_mels))
        d[0] /= np.sqrt(2.0)
        return d.astype(np.float32)


@dataclass
class MFCC:
    # Input
    signal: np.ndarray
    sample_rate: int | float

    # Settings
    n_mfcc: int = field(default=13)

    # Internals
    _feature_vector: np.ndarray = field(init=False)

    def __post_init__(self) -> None:
        _validate(self.signal)
        if self.n_mfcc != 13:
            raise NotImplementedError("the B200 MFCC kernel is built for n_mfcc=13 (the value every reference call site uses)")
        self._feature_vector = self.batch([self.signal], self.sample_rate)[0].T

    @property
    def feature_vector(self) -> np.ndarray:
        return self._feature_vector

    @staticmethod
    def normalize_mfccs(mfccs):
        """mfcc.py:50-69: statistics over axis 0, i.e. over the coefficients of each frame
        (host helper kept for API compatibility; the kernel applies the same formula)."""
        mean = np.mean(mfccs, axis=0, keepdims=True)
        std = np.std(mfccs, axis=0, keepdims=True)
        return (mfccs - mean) / (std + 1e-8)

    @classmethod
    def batch(cls, signals: List[NDArray], sample_rate: int, config: "MFCCConfig | None" = None) -> List[NDArray[np.float32]]:
        """List of (T, 39) float32 feature matrices (row = frame), one kernel pass for all signals.
        ``config`` (added, default = the reference's parameter set) selects another front end, see MFCCConfig."""
        from ._engine import get_engine

        for s in signals:
            _validate(s)
        if len(signals) == 0:
            return []
        eng = get_engine()
        b = eng.mfcc(signals, sample_rate, config)
        flat = b.feat.cpu().numpy()
        off = b.frm_off_host
        return [flat[off[i]:off[i + 1]] for i in range(len(signals))]


def _validate(signal) -> None:
    if not isinstance(signal, np.ndarray):
        raise TypeError("Input signal must be a numpy array.")
    if signal.ndim != 1:
        raise ValueError("Input signal must be 1-dimensional.")
