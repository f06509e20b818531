"""Oracle: vectorised NumPy restatement of the reference's HMM hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pinned bit-for-bit against the
unmodified reference (tests/golden/make_golden.py, tests/test_oracle_golden.py).

All citations are to /root/reference/src/loe_speech_recognition/.

A *trellis* is the flat description every Viterbi variant of the reference reduces to:

  band[p, k]  float32  log-transition into position p from position p-k (k = 0, 1, 2);
                       -inf where the reference never looks (p-k < 0, or below the word's
                       lower boundary in the loop grammar, hidden_markov_model.py:518)
  init[p]     bool     positions that receive  logpdf_p(x_0) + band[p, 0]  at t = 0
                       (state 0: :81-83;  every word start: :464-467)
  ends        int[]    termination candidates (last state: :198;  word ends: :566-571)
  loop        None, or (starts int[W], ends int[W], penalty, f64_mode): the word-start rule
                       of the digit-loop grammar (:533-559)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg

LOG_2PI = np.log(2 * np.pi)
NEG_INF = -np.inf


# --------------------------------------------------------------------------------------
# a2  Gaussian emission  (hidden_markov_model.py:27-48 -> scipy _multivariate.py _PSD/_logpdf)
# --------------------------------------------------------------------------------------
def gaussian_pack(mean, cov):
    """What ``scipy.stats.multivariate_normal(mean, cov, allow_singular=False)`` precomputes.

    Returns (mean f64 [D], U f64 [D,D], log_pdet f64).  Raises LinAlgError for a singular
    covariance and ValueError for a non-PSD one, like scipy's ``_PSD``.
    """
    mean = np.asarray(mean, dtype=np.float64)
    cov = np.asarray(cov, dtype=np.float64)
    s, u = scipy.linalg.eigh(cov, lower=True, check_finite=True)
    eps = 1e6 * np.finfo(s.dtype).eps * np.max(np.abs(s))      # _eigvalsh_to_eps, float64
    if np.min(s) < -eps:
        raise ValueError("The input matrix must be symmetric positive semidefinite.")
    if np.any(s <= eps):
        raise np.linalg.LinAlgError("singular covariance with allow_singular=False")
    U = u * np.sqrt(1.0 / s)
    return mean, U, float(np.sum(np.log(s)))


def emission_scores(x, means, Us, log_pdets):
    """[T, S] float32 matrix of ``MultivariateNormal.log_pdf`` (:46-48) values.

    x [T, D] float32; means [S, D] f64; Us [S, D, D] f64; log_pdets [S] f64.
    float64 arithmetic, one rounding to float32 at the end, as in the reference.
    """
    x = np.asarray(x)
    T, D = x.shape
    S = len(means)
    out = np.empty((T, S), dtype=np.float32)
    for s in range(S):
        dev = x - means[s]                       # float32 - float64 -> float64
        maha = np.sum(np.square(dev @ Us[s]), axis=-1)
        out[:, s] = (-0.5 * (D * LOG_2PI + log_pdets[s] + maha)).astype(np.float32)
    return out


# --------------------------------------------------------------------------------------
# trellis builders
# --------------------------------------------------------------------------------------
@dataclass
class Trellis:
    band: np.ndarray                 # [P, 3] float32
    init: np.ndarray                 # [P] bool
    ends: np.ndarray                 # [E] int
    loop_starts: Optional[np.ndarray] = None
    loop_ends: Optional[np.ndarray] = None

    @property
    def n_pos(self):
        return self.band.shape[0]


def _band_from_dense(logA, lower_of):
    """band[p,k] = logA[p-k, p] when p-k >= lower_of[p], else -inf."""
    P = logA.shape[0]
    band = np.full((P, 3), NEG_INF, dtype=np.float32)
    for p in range(P):
        for k in range(3):
            o = p - k
            if o >= lower_of[p]:
                band[p, k] = logA[o, p]
    return band


def word_trellis(logA):
    """Single word (hidden_markov_model.py:80-91, 160-208). logA: dense [S,S] float32."""
    S = logA.shape[0]
    init = np.zeros(S, dtype=bool)
    init[0] = True
    return Trellis(_band_from_dense(np.asarray(logA, np.float32), np.zeros(S, int)), init, np.array([S - 1]))


def block_diag_missing_zero(logAs):
    """LogTransitionProbabilities.append (transition_probability.py:70-75): block-diagonal
    concatenation where absent keys read back as 0.0 (:17-23)."""
    P = sum(a.shape[0] for a in logAs)
    out = np.zeros((P, P), dtype=np.float32)
    o = 0
    for a in logAs:
        n = a.shape[0]
        out[o:o + n, o:o + n] = a
        o += n
    return out


def chain_trellis(logAs):
    """Forced-alignment chain of embedded training (:638-664 + inherited _viterbi :80-91):
    the word Viterbi over the block-diagonal matrix, cross-word look-ups = 0.0."""
    return word_trellis(block_diag_missing_zero(logAs))


def boundaries(sizes):
    """ModelBoundary lower/upper boundaries (model_boundary.py:25-55)."""
    cum = np.cumsum(sizes)
    upper = cum - 1
    lower = np.concatenate(([0], cum[:-1]))
    return lower.astype(int), upper.astype(int)


def loop_trellis(logAs):
    """Digit-loop grammar (:463-581): within-word band restricted to the word, every word
    start initialised, termination over word ends, word-start rule handled by ``loop``."""
    sizes = [a.shape[0] for a in logAs]
    lower, upper = boundaries(sizes)
    P = int(sum(sizes))
    dense = block_diag_missing_zero(logAs)
    lower_of = np.repeat(lower, sizes)
    band = _band_from_dense(dense, lower_of)
    for lb in lower:                              # start states only use their self loop (:536-538)
        band[lb, 1:] = NEG_INF
    init = np.zeros(P, dtype=bool)
    init[lower] = True
    return Trellis(band, init, upper.copy(), lower.copy(), upper.copy())


def penalty_mode(penalty):
    """(value, f64_mode).  np.float64 penalties (the default np.log(0.005), :419) make the
    word-start candidates float64; Python scalars / np.float32 are weak -> float32."""
    f64 = isinstance(penalty, (np.float64,)) or (isinstance(penalty, np.ndarray) and penalty.dtype == np.float64)
    return penalty, bool(f64)


# --------------------------------------------------------------------------------------
# a3 / a4  Viterbi + backtrace
# --------------------------------------------------------------------------------------
def viterbi(scores, tr: Trellis, penalty=None):
    """Viterbi over one utterance given its emission scores.

    scores [T, P] float32 (column p = log_pdf of position p).  Returns
    (end_scores float32 [E], best_end_index, path int8 [T]) with the reference's exact
    arithmetic, tie-breaking (lowest index) and off-by-one backtrace (:201-207, :574-580).
    """
    scores = np.asarray(scores, dtype=np.float32)
    T, P = scores.shape
    band = tr.band
    ninf32 = np.float32(NEG_INF)
    d = np.full(P, ninf32, dtype=np.float32)
    d[tr.init] = scores[0, tr.init] + band[tr.init, 0]
    tracer = np.full((T, P), -1, dtype=np.int64)
    pos = np.arange(P)
    is_loop = tr.loop_starts is not None
    if is_loop:
        pen, f64_mode = penalty_mode(penalty)
        W = len(tr.loop_starts)
    for t in range(1, T):
        c0 = band[:, 0] + d
        c1 = np.full(P, ninf32, dtype=np.float32)
        c2 = np.full(P, ninf32, dtype=np.float32)
        c1[1:] = band[1:, 1] + d[:-1]
        c2[2:] = band[2:, 2] + d[:-2]
        best = c2.copy()
        arg = pos - 2
        m = c1 > best
        best[m] = c1[m]; arg[m] = pos[m] - 1
        m = c0 > best
        best[m] = c0[m]; arg[m] = pos[m]
        arg[best == ninf32] = 0                   # np.argmax of an all -inf array (:186, :523)
        new = (best.astype(np.float64) + scores[t].astype(np.float64)).astype(np.float32)
        if is_loop:
            dl = d[tr.loop_ends]
            if f64_mode:
                cand = np.float64(pen) + dl.astype(np.float64)
            else:
                cand = (np.float32(pen) + dl).astype(np.float64)
            k = int(np.argmax(cand))
            cbest = cand[k]
            for s in tr.loop_starts:
                self_c = np.float64(np.float32(band[s, 0] + d[s]))
                if self_c > cbest:                # self loop is the LAST array entry (:536, :546-550)
                    mv, bp = self_c, s
                else:
                    mv, bp = cbest, tr.loop_ends[k]
                new[s] = np.float32(mv + np.float64(scores[t, s]))
                arg[s] = bp
        tracer[t] = arg
        d = new
    end_scores = d[tr.ends].copy()
    bi = int(np.argmax(end_scores))
    prev = tracer[T - 1, tr.ends[bi]]
    path = np.zeros(T, dtype=np.int8)
    path[T - 1] = prev
    for t in range(T - 2, -1, -1):
        path[t] = prev
        prev = tracer[t, prev]
    return end_scores, bi, path


def viterbi_batch(scores_list, tr: Trellis, penalty=None):
    """Same as :func:`viterbi` for many utterances at once (vectorised over utterances so the
    full-size configurations finish in seconds).  Returns (end_scores [N,E], best [N], paths list)."""
    N = len(scores_list)
    lens = np.array([s.shape[0] for s in scores_list])
    P = tr.n_pos
    Tm = int(lens.max())
    band = tr.band
    ninf32 = np.float32(NEG_INF)
    sc = np.zeros((N, Tm, P), dtype=np.float32)
    for i, s in enumerate(scores_list):
        sc[i, : s.shape[0]] = s
    d = np.full((N, P), ninf32, dtype=np.float32)
    d[:, tr.init] = sc[:, 0][:, tr.init] + band[tr.init, 0]
    tracer = np.full((N, Tm, P), -1, dtype=np.int16)
    pos = np.arange(P)[None, :]
    is_loop = tr.loop_starts is not None
    if is_loop:
        pen, f64_mode = penalty_mode(penalty)
    rows = np.arange(N)
    final = np.full((N, P), ninf32, dtype=np.float32)
    final[lens == 1] = d[lens == 1]
    for t in range(1, Tm):
        c0 = band[None, :, 0] + d
        c1 = np.full((N, P), ninf32, dtype=np.float32)
        c2 = np.full((N, P), ninf32, dtype=np.float32)
        c1[:, 1:] = band[None, 1:, 1] + d[:, :-1]
        c2[:, 2:] = band[None, 2:, 2] + d[:, :-2]
        best = c2.copy()
        arg = np.broadcast_to(pos - 2, (N, P)).copy()
        m = c1 > best
        best[m] = c1[m]; arg[m] = np.broadcast_to(pos - 1, (N, P))[m]
        m = c0 > best
        best[m] = c0[m]; arg[m] = np.broadcast_to(pos, (N, P))[m]
        arg[best == ninf32] = 0
        new = (best.astype(np.float64) + sc[:, t].astype(np.float64)).astype(np.float32)
        if is_loop:
            dl = d[:, tr.loop_ends]
            if f64_mode:
                cand = np.float64(pen) + dl.astype(np.float64)
            else:
                cand = (np.float32(pen) + dl).astype(np.float64)
            k = np.argmax(cand, axis=1)
            cbest = cand[rows, k]
            for s in tr.loop_starts:
                self_c = (band[s, 0] + d[:, s]).astype(np.float32).astype(np.float64)
                take_self = self_c > cbest
                mv = np.where(take_self, self_c, cbest)
                new[:, s] = (mv + sc[:, t, s].astype(np.float64)).astype(np.float32)
                arg[:, s] = np.where(take_self, s, tr.loop_ends[k])
        tracer[:, t] = arg
        d = new
        done = lens == t + 1
        final[done] = d[done]
    end_scores = final[:, tr.ends]
    bi = np.argmax(end_scores, axis=1)
    paths = []
    for i in range(N):
        T = int(lens[i])
        prev = tracer[i, T - 1, tr.ends[bi[i]]]
        path = np.zeros(T, dtype=np.int8)
        path[T - 1] = prev
        for t in range(T - 2, -1, -1):
            path[t] = prev
            prev = tracer[i, t, prev]
        paths.append(path)
    return end_scores, bi, paths


def path_score(scores, tr: Trellis, states, penalty=None):
    """Re-score an explicit state sequence (true states s_0..s_{T-1}) with float64 arithmetic;
    used by the margin test that adjudicates path mismatches."""
    total = float(scores[0, states[0]]) + float(tr.band[states[0], 0])
    starts = set() if tr.loop_starts is None else set(int(s) for s in tr.loop_starts)
    ends = set() if tr.loop_ends is None else set(int(s) for s in tr.loop_ends)
    for t in range(1, len(states)):
        o, n = int(states[t - 1]), int(states[t])
        if n in starts and o != n:
            assert o in ends
            total += float(penalty)
        else:
            total += float(tr.band[n, n - o])
        total += float(scores[t, n])
    return total


# --------------------------------------------------------------------------------------
# a4  path -> label string  (model_boundary.py:107-147)
# --------------------------------------------------------------------------------------
def get_labels(path, sizes, labels, skip_silence=True):
    lower, upper = boundaries(sizes)
    path = [int(p) for p in path]

    def word_of(state):
        for w in range(len(lower) - 1, -1, -1):
            if state >= lower[w]:
                if state > upper[-1]:
                    raise Exception("state beyond the last word")
                return w
        raise Exception("state below the first word")     # model_boundary.py:68-70 (bare Exception)

    comp = [path[0]]
    for p in path[1:]:
        if p != comp[-1]:
            comp.append(p)
    out = []

    def emit(state):
        lab = labels[word_of(state)]
        if not (lab == "S" and skip_silence):
            out.append(lab)

    w = word_of(comp[0])
    emit(comp[0])
    for i in range(1, len(comp)):
        cur = comp[i]
        if cur < lower[w] or cur > upper[w]:
            w = word_of(cur)
            emit(cur)
        elif comp[i - 1] == upper[w] and cur == lower[w]:
            emit(cur)
    return out


# --------------------------------------------------------------------------------------
# a5  segmental K-means M-step  (hidden_markov_model.py:320-350, signal.py:23-47, 68-91)
# --------------------------------------------------------------------------------------
class TrainMeanFail(Exception):
    pass


def order_by_state(signal, path, n_states):
    """Signal.order_by_state (signal.py:23-47): contiguous run per state in increasing order."""
    segs = []
    start = 0
    for s in range(n_states):
        end = start
        while end < len(path) and path[end] == s:
            end += 1
        segs.append(signal[start:end] if start < end else None)
        start = end
    return segs


def mstep(signals, paths, n_states, old_means=None):
    """One M-step.  Returns dict(means f32 [S,D], covs f32 [S,D,D], trans f32 [S,S],
    counts int32 [S,S], occupancy int [S], converged bool).

    ``converged`` is the reference's np.allclose(new_means, old_means) test, evaluated
    BEFORE covariances / transitions are touched (:333-335)."""
    by_state = [[] for _ in range(n_states)]
    for sig, path in zip(signals, paths):
        for s, seg in enumerate(order_by_state(sig, path, n_states)):
            if seg is not None:
                by_state[s].append(seg)
    try:
        concat = [np.concatenate(b) for b in by_state]
    except ValueError:
        raise TrainMeanFail
    new_means = [np.average(c, axis=0) for c in concat]
    converged = old_means is not None and bool(np.allclose(new_means, old_means))
    means = np.array(new_means, dtype=np.float32)
    D = means.shape[1]
    covs = np.zeros((n_states, D, D), dtype=np.float32)
    with np.errstate(all="ignore"):
        for s, c in enumerate(concat):
            covs[s] = (np.cov(c, rowvar=False) + np.eye(D) * 0.001).astype(np.float32)
        counts = np.zeros((n_states, n_states), dtype=np.int32)
        for path in paths:
            last = path[0]
            for cur in path[1:]:
                counts[last, cur] += 1
                last = cur
        trans = (counts / np.sum(counts, axis=1, keepdims=True)).astype(np.float32)
    return dict(means=means, covs=covs, trans=trans, counts=counts,
                occupancy=np.array([len(c) for c in concat]), converged=converged)


def init_parameters(sample, n_states):
    """_init_parameters (:359-389): uniform slices of the FIRST utterance, cov = 0.01 I,
    transitions uniform over the current and all later states (transition_probability.py:42-52)."""
    D = sample.shape[1]
    L = int(sample.shape[0] / n_states)
    means = np.array([np.average(sample[i * L:(i + 1) * L], axis=0) for i in range(n_states)], dtype=np.float32)
    covs = (np.tile(np.eye(D), (n_states, 1, 1)) * 0.01).astype(np.float32)
    trans = np.zeros((n_states, n_states), dtype=np.float32)
    for i in range(n_states):
        trans[i, i:] = np.float32(1 / (n_states - i))
    return means, covs, trans


def log_transitions(trans):
    with np.errstate(divide="ignore"):
        return np.log(np.asarray(trans, dtype=np.float32))


# --------------------------------------------------------------------------------------
# a6  embedded training: chain labels and remux  (hidden_markov_model.py:602-636, 794-797)
# --------------------------------------------------------------------------------------
def insert_silence(labels: str) -> str:
    return "".join(f"S{c}" for c in labels) + "S"


def remux(signal, path, sizes, labels):
    """_remux_path_and_signal: cut the chain alignment where the word LABEL changes, re-base
    each piece to its word's first state; the final piece is never flushed (:614-636).
    Returns {label: [(segment, rebased_path, n_states), ...]}."""
    lower, upper = boundaries(sizes)

    def inst_of(state):
        for w in range(len(lower) - 1, -1, -1):
            if state >= lower[w]:
                return w
        raise Exception

    out: Dict[str, list] = {lab: [] for lab in labels}
    last_index = 0
    last_inst = inst_of(int(path[0]))
    for index in range(len(path)):
        inst = inst_of(int(path[index]))
        if labels[inst] != labels[last_inst]:
            lo = lower[last_inst]
            up_w = [w for w in range(len(upper)) if int(path[last_index]) <= upper[w]][0]
            n = upper[up_w] - lo + 1
            out[labels[last_inst]].append((signal[last_index:index], (path[last_index:index] - lo).astype(np.int8), int(n)))
            last_index = index
            last_inst = inst
    return out
