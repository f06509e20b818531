"""Oracle: the reference's time-synchronous template DTW.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Restates
``src/loe_speech_recognition/dynamic_time_wrapping.py:31-120`` on FEATURES (the reference computes
them with its MFCC class in ``__post_init__``; here they are passed in), column by column, including
its quirks:
  * row ``start_w`` of word w > 0 is also the last row of word w-1: it is computed twice per column,
    word w's value survives, but both values feed the column minimum used for pruning (:82-103);
  * row 0 of word 0 indexes template frame -1 and cost row -1 (Python wrap-around, :76, :82);
  * the reported distance of word w is read one row above its last frame (:106-107);
  * local distance = float32 sqrt(sum((a-b)^2)) (:118-120), accumulated costs float64.
Pinned against the real class in tests/test_dtw.py (authoring container) and tests/golden/golden_dtw.npz.
"""
from __future__ import annotations

import math

import numpy as np


def search(seq_feats, sample_feat, pruning=True, pruning_factor=4, trace_back=False):
    """seq_feats: list of (T_w, D) float32 template features; sample_feat (L, D) float32.
    Returns (index, min_distance, cost_matrix, path_matrix)."""
    lens = [f.shape[0] for f in seq_feats]
    seq = np.concatenate(seq_feats)
    H, L = seq.shape[0], sample_feat.shape[0]
    starts = [0]
    for n in lens[:-1]:
        starts.append(starts[-1] + n)
    cost = np.zeros((H + 1, L + 1))
    cost[1:, 0] = math.inf
    for p in starts:
        cost[p, 1:] = math.inf
        cost[p, 0] = 0
    path = np.zeros((H + 1, L + 1), dtype=int)
    min_col = np.full(L + 1, math.inf)
    for j in range(1, L + 1):
        min_col[j] = math.inf
        x = sample_feat[j - 1]
        for start, n in zip(starts, lens):
            for i in range(start, start + n + 1):
                d = np.sqrt(np.sum((seq[i - 1] - x) ** 2))
                ins = cost[i, j - 1]
                shr = math.inf if i - 2 < start else cost[i - 2, j - 1]
                mat = cost[i - 1, j - 1]
                m = min(ins, shr, mat)
                cur = d + m
                if pruning and cur > min_col[j - 1] * (1 + pruning_factor):
                    cost[i, j] = math.inf
                    continue
                cost[i, j] = cur
                if trace_back:
                    path[i, j] = 1 if m == ins else (2 if m == shr else 3)
                if cost[i, j] != math.inf:
                    min_col[j] = min(min_col[j], cost[i, j])
    dist = [cost[p + n - 1, L] for p, n in zip(starts, lens)]
    best = min(dist)
    return dist.index(best), best, cost, path
