"""Oracle: diagonal-covariance Gaussian-mixture emission scoring and the lexicon-expanded phone loop.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

PARITY UNPINNED BY CONSTRUCTION: the live reference scores ONE full-covariance Gaussian per state
(hidden_markov_model.py:20-48); its mixture code is abandoned and un-importable
(deprecated/gaussian_mixture_model.py).  That file fixes the semantics restated here:

  log p(x | state) = logaddexp over mixtures of  log w_m + log N(x; mu_m, diag var_m)     (:157-162)

BASELINE.json names the configurations: configs[0] extension set (11 words x 5 states, 4 Gaussians) and configs[4]
(~40 phones x 3 states, 16 Gaussians, lexicon-expanded digit loop).  The decode on top is the reference's own loop
grammar (hidden_markov_model.py:463-581, restated in oracle/hmm.py) with trellis positions mapped to shared phone states.
"""
from __future__ import annotations

import numpy as np

from . import hmm as O

LOG_2PI = np.log(2 * np.pi)


def component_constants(weights, variances):
    """c[s, m] = log w - 1/2 (D log 2pi + sum log var), float64."""
    weights = np.asarray(weights, dtype=np.float64)
    variances = np.asarray(variances, dtype=np.float64)
    D = variances.shape[-1]
    with np.errstate(divide="ignore"):
        return np.log(weights) - 0.5 * (D * LOG_2PI + np.sum(np.log(variances), axis=-1))


def gmm_emission_scores(x, weights, means, variances):
    """[T, S] float32: log sum_m w[s,m] N(x_t; means[s,m], diag variances[s,m]); float64 arithmetic, one rounding."""
    x = np.asarray(x, dtype=np.float64)
    means = np.asarray(means, dtype=np.float64)
    variances = np.asarray(variances, dtype=np.float64)
    c = component_constants(weights, variances)                       # [S, M]
    S, M, D = means.shape
    out = np.empty((x.shape[0], S), dtype=np.float32)
    for s in range(S):
        d = x[:, None, :] - means[s][None, :, :]                      # [T, M, D]
        comp = c[s][None, :] - 0.5 * np.sum(d * d / variances[s][None, :, :], axis=-1)
        mx = np.max(comp, axis=1, keepdims=True)
        with np.errstate(invalid="ignore"):
            lse = mx[:, 0] + np.log(np.sum(np.exp(comp - mx), axis=1))
        out[:, s] = np.where(np.isfinite(mx[:, 0]), lse, mx[:, 0]).astype(np.float32)
    return out


def word_log_transitions(phone_logA, phone_log_exit, phones):
    """Dense log-transition matrix of a word = chain of 3-state phones: phone blocks on the diagonal, the exit
    log-probability of a phone's last state on the entry into the next phone's first state, -inf elsewhere
    (stored zeros, like the reference's trained matrices: transition_probability.py:62-63)."""
    n = sum(phone_logA[p].shape[0] for p in phones)
    out = np.full((n, n), -np.inf, dtype=np.float32)
    o = 0
    for i, p in enumerate(phones):
        a = np.asarray(phone_logA[p], dtype=np.float32)
        k = a.shape[0]
        out[o:o + k, o:o + k] = a
        if i + 1 < len(phones):
            out[o + k - 1, o + k] = np.float32(phone_log_exit[p])
        o += k
    return out


def lexicon_trellis(phone_logA, phone_log_exit, phone_col, lexicon, order):
    """(Trellis, col int[P], sizes) of the word loop over ``order`` with every word expanded into its phones;
    col[p] = emission column (phone state) of trellis position p."""
    dense, cols = [], []
    for w in order:
        dense.append(word_log_transitions(phone_logA, phone_log_exit, lexicon[w]))
        cols.extend(phone_col[p] + j for p in lexicon[w] for j in range(phone_logA[p].shape[0]))
    return O.loop_trellis(dense), np.array(cols, dtype=np.int64), [d.shape[0] for d in dense]
