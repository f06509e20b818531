"""Oracle-side adjudication of decode mismatches (SURVEY.md §8d "margin test").

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): imported by tests/, __graft_entry__.smoke() and the
untimed correctness gate / cpu_baseline legs of bench.py -- never by the product.

A decoded state path that differs from the oracle's is excused only when BOTH state sequences, re-scored
with the oracle's own arithmetic (oracle MFCC -> oracle float32 emission scores -> float64 path score,
hidden_markov_model.py:463-581 of the reference), differ by at most ``rtol * |score|``: the two paths are
then a numerical near-tie that the 1e-4 feature / log-likelihood tolerance cannot resolve.  Everything
else is a failure and is reported, never absorbed into a percentage.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from . import hmm as O
from . import mfcc as OM


def true_states(path: np.ndarray, final_state: int) -> np.ndarray:
    """The reference's backtrace returns s_0 .. s_{T-2}, s_{T-2} (off by one, hidden_markov_model.py:201-207,
    574-580): the state sequence that was actually scored ends in ``final_state`` (the best END position)."""
    st = np.asarray(path, dtype=np.int64).copy()
    if len(st) > 1:
        st[-1] = int(final_state)
    return st


def final_state_of_loop_path(path: np.ndarray, sizes: Sequence[int]) -> int:
    """Final state of a loop-grammar path when the decoder did not report it: termination is over word-END states
    (hidden_markov_model.py:566-571), words have >= 2 states, so the last frame sits in the last state of the word
    that holds s_{T-2} (a word start entered at the last frame could not be an end state)."""
    lower, upper = O.boundaries(list(sizes))
    s = int(path[-1])
    w = int(np.searchsorted(upper, s))
    return int(upper[w])


@dataclass
class Verdict:
    n: int = 0
    identical_paths: int = 0
    identical_strings: int = 0
    excused: int = 0                       # path differs, both paths re-scored by the oracle within rtol
    failed: List[dict] = field(default_factory=list)
    worst_excused_rel: float = 0.0

    def summary(self) -> dict:
        return {"utterances": self.n, "paths_identical_vs_oracle": self.identical_paths,
                "strings_identical_vs_oracle": self.identical_strings, "margin_excused": self.excused,
                "margin_failed": len(self.failed), "worst_excused_rel_margin": self.worst_excused_rel}


def _path_allowed(tr: O.Trellis, states) -> bool:
    """Every step is a transition the trellis has: self / +1 / +2 inside the band, or word END -> word START."""
    P = tr.n_pos
    starts = set() if tr.loop_starts is None else {int(s) for s in tr.loop_starts}
    ends = set() if tr.loop_ends is None else {int(s) for s in tr.loop_ends}
    st = [int(s) for s in states]
    if any(s < 0 or s >= P for s in st) or not bool(tr.init[st[0]]) or st[-1] not in {int(e) for e in tr.ends}:
        return False
    for o, n in zip(st[:-1], st[1:]):
        if n in starts and o != n:
            if o not in ends:
                return False
        elif not (0 <= n - o <= 2) or not np.isfinite(tr.band[n, n - o]):
            return False
    return True


def adjudicate(scores_oracle, tr: O.Trellis, penalty, got_path, got_final, want_path, want_final, rtol=1e-4):
    """(ok, rel) for ONE utterance whose state paths differ: re-score both true state sequences on the oracle's
    float32 emission scores in float64 (O.path_score); ok when |delta| <= rtol * |oracle path score|.
    A path the grammar does not allow scores -inf / raises and is never excused."""
    ga, wa = true_states(got_path, got_final), true_states(want_path, want_final)
    if not (_path_allowed(tr, ga) and _path_allowed(tr, wa)):
        return False, float("inf")
    a = O.path_score(scores_oracle, tr, ga, penalty)
    b = O.path_score(scores_oracle, tr, wa, penalty)
    if not (np.isfinite(a) and np.isfinite(b)):
        return False, float("inf")
    rel = abs(a - b) / max(abs(b), 1e-30)
    return bool(rel <= rtol), float(rel)


def compare_loop_decodes(ems_oracle, tr: O.Trellis, penalty, sizes, labels, got_paths, got_strings=None,
                         got_finals=None, rtol=1e-4) -> Verdict:
    """Oracle decode of every utterance from its oracle emission scores, compared with the paths (and strings)
    another decoder produced.  ``got_finals``: the decoder's best END state per utterance (None: inferred)."""
    es, bi, want_paths = O.viterbi_batch(ems_oracle, tr, penalty=penalty)
    v = Verdict(n=len(ems_oracle))
    for i, (em, gp, wp) in enumerate(zip(ems_oracle, got_paths, want_paths)):
        want_str = "".join(O.get_labels(wp, sizes, list(labels))) if len(wp) > 1 else None
        same_path = np.array_equal(np.asarray(gp), wp)
        if got_strings is not None and want_str is not None and got_strings[i] == want_str:
            v.identical_strings += 1
        if same_path:
            v.identical_paths += 1
            continue
        gf = int(got_finals[i]) if got_finals is not None else final_state_of_loop_path(gp, sizes)
        ok, rel = adjudicate(em, tr, penalty, gp, gf, wp, int(tr.ends[bi[i]]), rtol)
        if ok:
            v.excused += 1
            v.worst_excused_rel = max(v.worst_excused_rel, rel)
        else:
            v.failed.append({"utt": i, "rel_margin": rel, "got": None if got_strings is None else got_strings[i], "want": want_str})
    return v


class OracleLoopDecoder:
    """Oracle pipeline for the digit-loop decode: PCM -> oracle MFCC -> oracle emission -> (scores per utterance).
    ``params`` maps word -> (means, covs, logA) in grammar order."""

    def __init__(self, params: dict, order: Sequence[str]):
        self.order = list(order)
        packs = [O.gaussian_pack(params[w][0][s], params[w][1][s]) for w in self.order for s in range(len(params[w][0]))]
        self.means = np.array([p[0] for p in packs])
        self.Us = np.array([p[1] for p in packs])
        self.lps = np.array([p[2] for p in packs])
        self.sizes = [len(params[w][0]) for w in self.order]
        self.trellis = O.loop_trellis([params[w][2] for w in self.order])

    def features(self, utts):
        return [OM.mfcc_feature_vector(np.asarray(u, dtype=np.float32)).T for u in utts]

    def emissions(self, feats):
        return [O.emission_scores(x, self.means, self.Us, self.lps) for x in feats]

    def compare(self, utts, penalty, got_paths, got_strings=None, got_finals=None, rtol=1e-4) -> Verdict:
        ems = self.emissions(self.features(utts))
        return compare_loop_decodes(ems, self.trellis, penalty, self.sizes, self.order, got_paths, got_strings, got_finals, rtol)
