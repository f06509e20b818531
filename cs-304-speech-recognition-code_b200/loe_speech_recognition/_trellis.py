"""Host-side flattening of the reference's model objects into the flat tables the kernels read.

The reference keeps transitions in dict-backed sparse matrices (transition_probability.py:11-82)
and word boundaries in ModelBoundary (model_boundary.py:11-179) and queries them per
(frame, state) from Python.  Here they are flattened ONCE per model into:

  col[p]      emission column of trellis position p
  band[p, k]  log-transition into p from p-k, k = 0..2 (-inf = the reference never looks there)
  flags[p]    INIT / START / END bits
  word[p]     label id of the word instance p belongs to, word_lo[p] its first position

Three shapes (include/loe_b200.h, loe_viterbi_dev):
  word   - one left-to-right word              hidden_markov_model.py:80-91, 160-208
  multi  - several independent words side by side (isolated-word classifier, one launch
           instead of ModelCollection.predict's loop, model_collection.py:23-28)
  loop   - digit-loop grammar                  hidden_markov_model.py:463-581
  chain  - forced-alignment chain              hidden_markov_model.py:638-664 (missing keys = 0.0,
           transition_probability.py:17-23)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

from ._native import POS_END, POS_INIT, POS_START

NEG_INF = np.float32(-np.inf)


@dataclass
class HostTrellis:
    col: np.ndarray        # int32 [P]
    band: np.ndarray       # float32 [P, 3]
    flags: np.ndarray      # uint8 [P]
    word: np.ndarray       # int32 [P]
    word_lo: np.ndarray    # int32 [P]

    @property
    def n_pos(self) -> int:
        return int(self.col.shape[0])

    @property
    def n_ends(self) -> int:
        return int(np.count_nonzero(self.flags & POS_END))


def _band(dense: np.ndarray, lower_of: np.ndarray) -> np.ndarray:
    P = dense.shape[0]
    band = np.full((P, 3), NEG_INF, dtype=np.float32)
    idx = np.arange(P)
    for k in range(3):
        o = idx - k
        ok = o >= lower_of
        band[ok, k] = dense[o[ok], idx[ok]]
    return band


def _concat(dense_list: Sequence[np.ndarray], cross_value: float) -> np.ndarray:
    P = int(sum(d.shape[0] for d in dense_list))
    out = np.full((P, P), np.float32(cross_value), dtype=np.float32)
    o = 0
    for d in dense_list:
        n = d.shape[0]
        out[o:o + n, o:o + n] = d
        o += n
    return out


def build(dense_list: Sequence[np.ndarray], cols: Sequence[int], label_ids: Sequence[int], kind: str) -> HostTrellis:
    """dense_list[i]: dense [S_i, S_i] float32 log-transition matrix of word instance i;
    cols[i]: emission column of its first state; label_ids[i]: id of its label."""
    sizes = np.array([d.shape[0] for d in dense_list], dtype=np.int64)
    P = int(sizes.sum())
    lower = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int64)
    upper = lower + sizes - 1
    lower_of = np.repeat(lower, sizes)
    col = np.concatenate([np.arange(c, c + n) for c, n in zip(cols, sizes)]).astype(np.int32)
    word = np.repeat(np.asarray(label_ids, dtype=np.int32), sizes)
    flags = np.zeros(P, dtype=np.uint8)
    if kind in ("word", "chain"):
        dense = _concat(dense_list, 0.0)          # absent keys read back as 0.0
        band = _band(dense, np.zeros(P, dtype=np.int64))
        flags[0] |= POS_INIT
        flags[P - 1] |= POS_END
    elif kind in ("multi", "loop"):
        dense = _concat(dense_list, 0.0)
        band = _band(dense, lower_of)             # never below the word's own first state
        flags[lower] |= POS_INIT
        flags[upper] |= POS_END
        if kind == "loop":
            flags[lower] |= POS_START
            band[lower, 1:] = NEG_INF             # word starts: self loop or word ends only (:533-559)
    else:
        raise ValueError(kind)
    return HostTrellis(col, band, flags, word, lower_of.astype(np.int32))


def stack(trellises: List[HostTrellis]):
    """Concatenate trellises into the table layout of the C ABI. Returns
    (tr_off int32 [n+1], col, band [P,3], flags, word, word_lo, max_pos, max_ends)."""
    off = np.concatenate(([0], np.cumsum([t.n_pos for t in trellises]))).astype(np.int32)
    return (off,
            np.concatenate([t.col for t in trellises]).astype(np.int32),
            np.ascontiguousarray(np.concatenate([t.band for t in trellises]), dtype=np.float32),
            np.concatenate([t.flags for t in trellises]).astype(np.uint8),
            np.concatenate([t.word for t in trellises]).astype(np.int32),
            np.concatenate([t.word_lo for t in trellises]).astype(np.int32),
            int(max(t.n_pos for t in trellises)),
            int(max(t.n_ends for t in trellises)))
