"""Seeded synthetic corpus for parity tests and benchmarks.

The reference ships no data (TIDIGITS is not redistributable, README.md:5), so every
measurement uses waveforms generated here: per (word, state) a fixed set of 2-3 sinusoids
in 200-4000 Hz with amplitude ~3000 plus white noise (sigma ~30); silence is noise only with a
state-dependent level.  Samples are float32 at int16 scale, integer valued, exactly what
``scipy.io.wavfile.read(...).astype(np.float32)`` hands the reference (ti_digits.py:133).

This module is additive (the reference has no counterpart) and host-only.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

SAMPLE_RATE = 16000
DIGITS: Tuple[str, ...] = ("1", "2", "3", "4", "5", "6", "7", "8", "9", "O", "Z")
SILENCE = "S"
WORDS: Tuple[str, ...] = DIGITS + (SILENCE,)
STATES_PER_WORD: Dict[str, int] = {**{d: 5 for d in DIGITS}, SILENCE: 3}
_SIL_SIGMA = (18.0, 30.0, 48.0)


def _tones(word: str):
    """Deterministic (independent of the corpus seed) tone table of one word."""
    rng = np.random.default_rng(1000 + WORDS.index(word))
    out = []
    for _ in range(STATES_PER_WORD[word]):
        k = int(rng.integers(2, 4))
        out.append((rng.uniform(200.0, 4000.0, size=k), rng.uniform(2000.0, 4000.0, size=k)))
    return out


_TONES = {w: _tones(w) for w in DIGITS}


def synth_word(rng: np.random.Generator, word: str, n_samples: int) -> Tuple[np.ndarray, np.ndarray]:
    """One word of ``n_samples`` samples with near-uniform state durations.
    Returns (pcm float32, per-sample state index int8)."""
    n_states = STATES_PER_WORD[word]
    edges = np.linspace(0, n_samples, n_states + 1)
    jitter = rng.uniform(-0.08, 0.08, size=n_states - 1) * (n_samples / n_states)
    edges[1:-1] += jitter
    edges = np.round(edges).astype(int)
    t = np.arange(n_samples) / SAMPLE_RATE
    pcm = np.zeros(n_samples, dtype=np.float64)
    state = np.zeros(n_samples, dtype=np.int8)
    for s in range(n_states):
        a, b = edges[s], edges[s + 1]
        state[a:b] = s
        if word == SILENCE:
            pcm[a:b] = rng.normal(0.0, _SIL_SIGMA[s], size=b - a)
        else:
            freqs, amps = _TONES[word][s]
            seg = np.zeros(b - a)
            for f, amp in zip(freqs, amps):
                seg += amp * np.sin(2 * np.pi * f * t[a:b] + rng.uniform(0, 2 * np.pi))
            pcm[a:b] = seg + rng.normal(0.0, 30.0, size=b - a)
    pcm = np.clip(np.round(pcm), -32767, 32767).astype(np.float32)
    return pcm, state


def synth_isolated(rng, word: str, seconds: float = None) -> np.ndarray:
    """An isolated word (config 1 / 3).  ``seconds`` None -> the same duration law as inside
    a digit string (digits 0.30-0.45 s, silence 0.10-0.20 s), so that isolated-word models
    align the strings of config 2."""
    if seconds is None:
        seconds = float(rng.uniform(0.10, 0.20) if word == SILENCE else rng.uniform(0.30, 0.45))
    pcm, _ = synth_word(rng, word, int(round(seconds * SAMPLE_RATE)))
    return pcm


def synth_string(rng, digits: Sequence[str], return_segments: bool = False):
    """S + n x (digit + S): digits 0.30-0.45 s, silences 0.10-0.20 s (config 2).
    With ``return_segments`` also returns [(word, first_sample, end_sample), ...]."""
    words = [SILENCE]
    for d in digits:
        words += [d, SILENCE]
    parts, segs, pos = [], [], 0
    for w in words:
        lo, hi = (0.10, 0.20) if w == SILENCE else (0.30, 0.45)
        n = int(rng.uniform(lo, hi) * SAMPLE_RATE)
        parts.append(synth_word(rng, w, n)[0])
        segs.append((w, pos, pos + n))
        pos += n
    pcm = np.concatenate(parts)
    return (pcm, segs) if return_segments else pcm


def silence_frames(features: np.ndarray, segments, hop: int = 160, min_frames: int = 6) -> List[np.ndarray]:
    """Slices of a (T, D) feature matrix that lie inside the silence segments of a string.
    The front end references every frame to the utterance maximum (power_to_db(ref=np.max),
    mfcc.py:35), so a silence model has to be trained on in-context silence -- the reference
    does the same by collecting noise with SignalSeparation (scripts/project5_train_no_empty.py)."""
    out = []
    for w, a, b in segments:
        if w != SILENCE:
            continue
        fa, fb = a // hop + 2, b // hop - 1
        if fb - fa >= min_frames:
            out.append(features[fa:fb])
    return out


def isolated_corpus(seed: int, n_per_word: int, words: Sequence[str] = WORDS, seconds=None) -> Dict[str, List[np.ndarray]]:
    rng = np.random.default_rng(seed)
    return {w: [synth_isolated(rng, w, seconds) for _ in range(n_per_word)]
            for w in words}


def string_corpus(seed: int, n_utts: int, n_digits: int = 7) -> Tuple[List[np.ndarray], List[str]]:
    rng = np.random.default_rng(seed)
    utts, truth = [], []
    for _ in range(n_utts):
        ds = [DIGITS[int(i)] for i in rng.integers(0, len(DIGITS), size=n_digits)]
        utts.append(synth_string(rng, ds))
        truth.append("".join(ds))
    return utts, truth
