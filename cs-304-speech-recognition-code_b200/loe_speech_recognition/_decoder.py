"""Torch-free binding of the host-buffer decoder (``loe_decoder_*`` in include/loe_b200.h):
ctypes + numpy only.  The C side owns device memory, streams and the chunked copy/compute
overlap; this module only prepares the model tables (the same host packs the torch-backed engine
uploads) and turns word-id tables into strings.

It replaces, for a batch, ``[inference.predict(MFCC(sig, sr).feature_vector) for sig in signals]``
(reference mfcc.py:24-44, hidden_markov_model.py:458-581, model_boundary.py:107-147).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native


def _ptr(a: Optional[np.ndarray]) -> int:
    return 0 if a is None else a.ctypes.data


class PinnedBuffer:
    """Page-locked host array (``loe_host_alloc``): PCM placed here is copied asynchronously, so the
    copy of one chunk overlaps the kernels of the previous one."""

    def __init__(self, n: int, dtype=np.float32):
        self._lib = _native.load()
        self.dtype = np.dtype(dtype)
        self.nbytes = int(n) * self.dtype.itemsize
        p = ctypes.c_void_p()
        _native.check(self._lib.loe_host_alloc(ctypes.byref(p), self.nbytes))
        self._p = p
        buf = (ctypes.c_char * max(self.nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(n))

    def close(self) -> None:
        if self._p is not None:
            self.array = None
            self._lib.loe_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NativeDecoder:
    """PCM (host) -> word ids (host) through ``loe_decoder_decode_host``.

    ``means / us / cst``: float64 Gaussian arrays of all states ([S,39], [S,39,39], [S]);
    ``trellis``: the loop-grammar HostTrellis (``_trellis.build(..., "loop")``)."""

    def __init__(self, means: np.ndarray, us: np.ndarray, cst: np.ndarray, trellis, sample_rate: float = 16000,
                 device: int = 0):
        from ._engine import default_precision, pack_h16_image, pack_tc_image
        from .mfcc import mel_lane_tables

        if means.shape[1] != 39:
            raise NotImplementedError("the host-buffer decoder is built for the 39-dimensional MFCC front end")
        self._lib = _native.load()
        bins, w, na, nb = mel_lane_tables(sample_rate)
        b_packed, cst_pad = pack_tc_image(means, us, cst)
        col = np.ascontiguousarray(trellis.col, dtype=np.int32)
        band = np.ascontiguousarray(trellis.band, dtype=np.float32)
        flags = np.ascontiguousarray(trellis.flags, dtype=np.uint8)
        word = np.ascontiguousarray(trellis.word, dtype=np.int32)
        word_lo = np.ascontiguousarray(trellis.word_lo, dtype=np.int32)
        handle = ctypes.c_void_p()
        _native.check(self._lib.loe_decoder_create(
            int(device), _ptr(bins), _ptr(w), na, nb, _ptr(b_packed), _ptr(cst_pad), int(means.shape[0]),
            int(col.shape[0]), _ptr(col), _ptr(band), _ptr(flags), _ptr(word), _ptr(word_lo), ctypes.byref(handle)))
        self._h = handle
        # same emission policy as the engine ("auto": 3xFP16 when the whitening matrices fit binary16)
        b_h16 = pack_h16_image(means, us, cst) if default_precision() in ("auto", "h16") else None
        if b_h16 is not None:
            b_h16 = np.ascontiguousarray(b_h16)
            _native.check(self._lib.loe_decoder_set_h16(self._h, _ptr(b_h16)))
        self.emission = "h16" if b_h16 is not None else "tc"
        self.b_h16 = b_h16
        self.sample_rate = sample_rate
        self.tables = (bins, w, na, nb, b_packed, cst_pad, int(means.shape[0]), col, band, flags, word, word_lo)

    def narrow_rate(self) -> float:
        """GB/s (float32 bytes) the library measured for its host-side float32 -> int16 narrowing: > 0 in use,
        < 0 measured and switched off, 0 not measured / disabled (include/loe_b200.h)."""
        return float(self._lib.loe_decoder_narrow_rate(self._h))

    def set_narrow(self, mode) -> None:
        """``"on"`` / ``"off"``: narrow float32 PCM to int16 on the host (or never); ``"auto"``: let the decoder decide
        from its own per-chunk measurements (include/loe_b200.h, loe_decoder_set_narrow)."""
        code = {"auto": -1, "off": 0, "on": 1, -1: -1, 0: 0, 1: 1, True: 1, False: 0}[mode]
        _native.check(self._lib.loe_decoder_set_narrow(self._h, code))

    def stats(self) -> dict:
        """What the last ``decode`` call did on the ingestion side (loe_decoder_stats)."""
        v = np.zeros(10, dtype=np.float64)
        _native.check(self._lib.loe_decoder_stats(self._h, _ptr(v), 10))
        return {"narrow_mode": {-1: "auto", 0: "off", 1: "on"}[int(v[0])],
                "narrow_on": {1: True, 0: False}.get(int(v[1])),          # None: still sampling
                "narrow_gbps": float(v[2]), "copy_gbps": float(v[3]), "narrow_threads": int(v[4]), "pinned_cpus": int(v[5]),
                "pcm_bytes": int(v[6]), "wire_bytes": int(v[7]), "chunks": int(v[8]), "chunks_narrowed": int(v[9])}

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            self._lib.loe_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode(self, pcm_flat: np.ndarray, sample_offsets: Sequence[int], penalty: float, penalty_f64: bool,
               skip_label: int = -1, max_words: int = 32, n_chunks: int = 0, want_scores: bool = False,
               want_path: bool = False) -> Tuple[np.ndarray, np.ndarray, Optional[np.ndarray], Optional[np.ndarray]]:
        """(words int8 [n, max_words], count int32 [n], best_score float32 [n] | None, path int8 [F] | None)."""
        if self._h is None:
            raise RuntimeError("decoder is closed")
        off = np.ascontiguousarray(sample_offsets, dtype=np.int64)
        n = off.shape[0] - 1
        pcm = np.asarray(pcm_flat)
        if pcm.dtype != np.int16 and pcm.dtype != np.float32:
            pcm = pcm.astype(np.float32)
        if not pcm.flags.c_contiguous:
            pcm = np.ascontiguousarray(pcm)
        if n > 0 and (off[0] != 0 or off[-1] > pcm.shape[0] or np.any(np.diff(off) < 0)):
            raise ValueError("sample_offsets must start at 0, be non-decreasing and end inside pcm_flat")
        words = np.empty((max(n, 0), max_words), dtype=np.int8)
        count = np.empty((max(n, 0),), dtype=np.int32)
        score = np.empty((n,), dtype=np.float32) if want_scores and n > 0 else None
        path = None
        if want_path and n > 0:
            path = np.empty((int((1 + np.diff(off) // 160).sum()),), dtype=np.int8)
        if n > 0:
            _native.check(self._lib.loe_decoder_decode_host(
                self._h, _ptr(pcm), 1 if pcm.dtype == np.int16 else 0, _ptr(off), n, float(penalty), 1 if penalty_f64 else 0,
                int(skip_label), int(max_words), int(n_chunks), _ptr(words), _ptr(count), _ptr(score), _ptr(path)))
        return words, count, score, path


def write_blob(path: str, decoder: NativeDecoder, pcm_flat: np.ndarray, sample_offsets, penalty: float, penalty_f64: bool,
               skip_label: int = -1, max_words: int = 32, n_chunks: int = 0) -> None:
    """Model tables + one PCM batch in the flat file examples/decode_host.c reads (a C caller's view
    of the same call :meth:`NativeDecoder.decode` makes)."""
    bins, w, na, nb, b_packed, cst_pad, n_states, col, band, flags, word, word_lo = decoder.tables
    off = np.ascontiguousarray(sample_offsets, dtype=np.int64)
    pcm = np.ascontiguousarray(pcm_flat)
    if pcm.dtype != np.int16:
        pcm = pcm.astype(np.float32)
    b_h16 = decoder.b_h16 if decoder.b_h16 is not None else np.zeros(0, dtype=np.float16)
    head = np.array([na, nb, n_states, col.shape[0], off.shape[0] - 1, max_words, 1 if pcm.dtype == np.int16 else 0,
                     1 if penalty_f64 else 0, skip_label, n_chunks, pcm.shape[0], b_h16.nbytes], dtype=np.int64)
    with open(path, "wb") as f:
        f.write(head.tobytes())
        f.write(np.array([penalty], dtype=np.float64).tobytes())
        for a in (bins, w, b_packed, cst_pad, col, band, flags, word, word_lo, b_h16, off, pcm):
            f.write(np.ascontiguousarray(a).tobytes())
