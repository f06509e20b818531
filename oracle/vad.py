"""Oracle: the reference's energy-hysteresis silence stripper.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Restates
``src/loe_speech_recognition/signal_separation.py:88-164`` as a pure function per signal; the
energies are computed with NumPy itself (``np.average(np.abs(frame))`` on float32, :155-157), so the
float32 pairwise summation order is the reference's by construction.  Pinned against the real class
in tests/test_vad.py (authoring container) and by tests/golden/golden_vad.npz.
"""
from __future__ import annotations

import numpy as np


def frame_energies(signal, frame_size):
    """float32 mean |x| of every 10 ms frame plus the trailing partial frame (may be empty -> nan)."""
    nf = signal.shape[0] // frame_size
    frames = list(signal[: frame_size * nf].reshape((-1, frame_size))) + [signal[frame_size * nf:]]
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return np.array([np.average(np.abs(f)) for f in frames], dtype=np.float32)


def segment(signal, sample_rate=16000, frame_time=0.01, high=0.08, low=0.01, silence_duration=0.02):
    """Returns dict(done, start, end, noise_mask, energies, max_volume):
    result frames are [start, end) (the frame that trips the silence counter is not included);
    noise_mask[f] marks frames the reference appends to its noise list; done=False means the
    reference raises FailToProcess because the speech never ended (:100-101)."""
    frame_size = int(sample_rate * frame_time)
    max_frames = int(silence_duration / frame_time)
    max_volume = np.max(np.abs(signal)).astype(float)
    hi_thr, lo_thr = high * max_volume, low * max_volume
    en = frame_energies(signal, frame_size)
    noise = np.zeros(len(en), dtype=bool)
    between = ever = False
    counter = 0
    start = -1
    done, end = False, len(en)
    for f, e in enumerate(en):
        tripped = False
        if between:
            if e > lo_thr:
                counter = 0
            else:
                between = False
                counter += 1
                tripped = counter >= max_frames
        else:
            if e > hi_thr:
                between = ever = True
                counter = 0
                if start < 0:
                    start = f
            else:
                noise[f] = True
                if ever:
                    counter += 1
                    tripped = counter >= max_frames
        if tripped:
            done, end = True, f
            break
    if start < 0:
        start = end
    return dict(done=done, start=start, end=end, noise_mask=noise, energies=en, max_volume=max_volume)


class Stripper:
    """Stateful wrapper with the reference's list semantics (noise carried over failed signals)."""

    def __init__(self, **kw):
        self.kw = kw
        self.frame_size = int(kw.get("sample_rate", 16000) * kw.get("frame_time", 0.01))
        self.noises = []
        self._noise = []

    def remove_empty(self, signal):
        """Returns the stripped signal or None where the reference raises FailToProcess."""
        r = segment(signal, **self.kw)
        fs = self.frame_size
        upto = r["end"] + 1 if r["done"] else len(r["energies"])
        for f in np.nonzero(r["noise_mask"][:upto])[0]:
            self._noise.append(signal[f * fs:(f + 1) * fs])
        if not r["done"]:
            return None
        self.noises.append(np.concatenate(self._noise, dtype=np.float32))
        self._noise = []
        if r["end"] - r["start"] < 9:
            return None
        return np.ascontiguousarray(signal[r["start"] * fs: r["end"] * fs], dtype=np.float32)
