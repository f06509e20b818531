"""Structure-faithful CPU port of the reference's decode path (the timed CPU baseline).

TEST / BASELINE INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``bench.py`` times this as
the reference's CPU implementation (``cpu_baseline.kind == "port"``, and the whole
``--impl reference`` arm): the reference itself is Python under /root/reference, which does
not exist on the GPU box.

Unlike ``oracle/hmm.py`` (vectorised, for checking), this file keeps the COST STRUCTURE of the
reference: one ``scipy.stats`` frozen ``logpdf`` call per (frame, state)
(hidden_markov_model.py:46-48, 189, 526, 556), a fresh float64 scratch array plus
``np.max``/``np.argmax`` per trellis cell (:180-186, :517-523, :534-547), dict-backed
transition look-ups (transition_probability.py:17-23) and linear boundary scans
(model_boundary.py:69-90).  It is pinned against the real reference in
tests/test_oracle_vs_reference.py (authoring container) and against oracle/hmm.py everywhere.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import scipy.stats


class SparseLog:
    """dict-backed log-transition matrix, absent key -> 0.0 (transition_probability.py:17-23)."""

    def __init__(self, blocks: Sequence[np.ndarray]):
        self.core: Dict[Tuple[int, int], np.float32] = {}
        base = 0
        for b in blocks:
            n = b.shape[0]
            for i in range(n):
                for j in range(n):
                    self.core[(i + base, j + base)] = np.float32(b[i, j])
            base += n
        self.n = base

    def __getitem__(self, key):
        assert any(key) < self.n
        if key in self.core:
            return self.core[key]
        return 0.0


class PortModel:
    """The state a reference ``HiddenMarkovModelInference`` holds after ``from_folder`` (:421-456)."""

    def __init__(self, means, covs, logAs, labels, penalty):
        self.normals = [scipy.stats.multivariate_normal(mean=m, cov=c, allow_singular=False)
                        for ms, cs in zip(means, covs) for m, c in zip(ms, cs)]
        self.logA = SparseLog(logAs)
        sizes = [a.shape[0] for a in logAs]
        cum = np.cumsum(sizes).tolist()
        self.lower = [0] + cum[:-1]
        self.upper = [c - 1 for c in cum]
        self.labels = list(labels)
        self.penalty = penalty

    def log_pdf(self, s, x):
        return self.normals[s].logpdf(x).astype(np.float32)

    def find_lower(self, state):
        for lb in reversed(self.lower):
            if state >= lb:
                return lb
        raise Exception

    def find_upper(self, state):
        for ub in self.upper:
            if state <= ub:
                return ub
        raise Exception


def loop_viterbi(model: PortModel, obs: np.ndarray):
    """HiddenMarkovModelInference._viterbi + _viterbi_static (:463-581), cell by cell."""
    n_states = len(model.normals)
    T = obs.shape[0]
    left = np.full((n_states,), -float("inf"), dtype=np.float32)
    for lb in model.lower:
        left[lb] = model.log_pdf(lb, obs[0]) + model.logA[lb, lb]
    right = np.full((n_states,), -float("inf"), dtype=np.float32)
    tracer = np.zeros((T, n_states), dtype=np.int8) - 1
    n_words = len(model.lower)
    for t in range(1, T):
        for new in range(n_states):
            if new in model.lower:
                continue
            lb = model.find_lower(new)
            cand = np.full((n_states,), -float("inf"))
            for old in range(max(new - 2, lb), new + 1):
                cand[old] = model.logA[(old, new)] + left[old]
            best = np.max(cand)
            arg = int(np.argmax(cand))
            right[new] = best + model.log_pdf(new, obs[t])
            tracer[t, new] = arg
        for new in model.lower:
            cand = np.full((n_words + 1,), -float("inf"))
            cand[-1] = model.logA[(new, new)] + left[new]
            for k, old in enumerate(model.upper):
                cand[k] = model.penalty + left[old]
            best = np.max(cand)
            k = int(np.argmax(cand))
            arg = new if k == n_words else model.upper[k]
            right[new] = best + model.log_pdf(new, obs[t])
            tracer[t, new] = arg
        left = right
        right = np.full((n_states,), -float("inf"), dtype=np.float32)
    ends = left[model.upper]
    best_score = np.max(ends)
    end_state = model.upper[int(np.argmax(ends))]
    prev = tracer[-1, end_state]
    path = np.zeros((T,), dtype=np.int8)
    path[-1] = prev
    for t in range(T - 2, -1, -1):
        path[t] = prev
        prev = tracer[t, prev]
    return best_score, path


def labels_from_path(model: PortModel, path: np.ndarray) -> str:
    """ModelBoundary.get_labels (model_boundary.py:107-147), silence dropped."""
    seq = path.tolist()
    comp = [seq[0]]
    for s in seq[1:]:
        if s != comp[-1]:
            comp.append(s)
    out: List[str] = []

    def emit(state):
        lab = model.labels[model.lower.index(model.find_lower(state))]
        if lab != "S":
            out.append(lab)

    lo, hi = model.find_lower(comp[0]), model.find_upper(comp[0])
    emit(comp[0])
    for i in range(1, len(comp)):
        cur = comp[i]
        if cur < lo or cur > hi:
            lo, hi = model.find_lower(cur), model.find_upper(cur)
            emit(cur)
        elif comp[i - 1] == hi and cur == lo:
            emit(cur)
    return "".join(out)


def decode_pcm(model: PortModel, pcm: np.ndarray) -> str:
    """One utterance end to end the way scripts/project5_test_ndigits_with_sil.py does it:
    MFCC (restated librosa, oracle/mfcc.py) -> loop Viterbi -> label string."""
    from . import mfcc as OM
    feats = OM.mfcc_feature_vector(pcm).T
    _, path = loop_viterbi(model, feats)
    return labels_from_path(model, path)


# ---- process-pool driver (the reference's own parallel form: ProcessPoolExecutor over utterances,
# scripts/project5_test_ndigits_with_sil.py:33-41) -------------------------------------------------
_POOL_MODEL = None


def _pool_init(args):
    global _POOL_MODEL
    _POOL_MODEL = PortModel(*args)


def _pool_decode(pcm):
    return decode_pcm(_POOL_MODEL, pcm)


def decode_pool(model_args, utterances: Sequence[np.ndarray], workers: int) -> List[str]:
    import concurrent.futures as cf
    import multiprocessing as mp
    if workers <= 1:
        _pool_init(model_args)
        return [_pool_decode(u) for u in utterances]
    with cf.ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("fork"),
                                initializer=_pool_init, initargs=(model_args,)) as ex:
        return list(ex.map(_pool_decode, utterances))
