"""CPU: the arbiter of the parity tests is itself tested -- oracle/adjudicate.py must count identical paths, excuse
exact ties only, and FAIL legal-but-worse paths and paths the grammar does not allow (the round-1 margin test could
never fail; SURVEY.md section 8d asks for both paths re-scored with the oracle's arithmetic)."""
import numpy as np
import pytest

from oracle import hmm as O
from oracle.adjudicate import adjudicate, compare_loop_decodes, final_state_of_loop_path, true_states

ORDER = ("1", "2", "3", "4", "5", "6", "7", "8", "9", "O", "S", "Z")


@pytest.fixture(scope="module")
def loop(golden):
    logAs = [golden[f"train_logA_{w}"] for w in ORDER]
    sizes = [len(a) for a in logAs]
    return O.loop_trellis(logAs), sizes


def _scores(rng, n, T, S):
    return [rng.normal(-60.0, 8.0, size=(T, S)).astype(np.float32) for _ in range(n)]


def test_identical_paths_are_counted_not_excused(loop):
    tr, sizes = loop
    ems = _scores(np.random.default_rng(0), 6, 120, sum(sizes))
    _, _, want = O.viterbi_batch(ems, tr, penalty=-100)
    strings = ["".join(O.get_labels(p, sizes, list(ORDER))) for p in want]
    v = compare_loop_decodes(ems, tr, -100, sizes, ORDER, want, strings)
    assert (v.n, v.identical_paths, v.identical_strings, v.excused, len(v.failed)) == (6, 6, 6, 0, 0)


def test_legal_but_worse_path_fails(loop):
    """The best path of DIFFERENT scores is a legal path of the grammar, but on the oracle's scores it is worse than
    the oracle's own path by far more than 1e-4 |score|: it must be reported, with its margin."""
    tr, sizes = loop
    rng = np.random.default_rng(1)
    ems = _scores(rng, 4, 150, sum(sizes))
    other = [e + rng.normal(0, 8.0, size=e.shape).astype(np.float32) for e in ems]
    _, _, got = O.viterbi_batch(other, tr, penalty=-100)
    v = compare_loop_decodes(ems, tr, -100, sizes, ORDER, got)
    assert v.identical_paths == 0 and v.excused == 0 and len(v.failed) == 4
    assert all(np.isfinite(f["rel_margin"]) and f["rel_margin"] > 1e-4 for f in v.failed)


def test_path_outside_the_grammar_is_never_excused(loop):
    tr, sizes = loop
    ems = _scores(np.random.default_rng(2), 1, 100, sum(sizes))
    _, bi, want = O.viterbi_batch(ems, tr, penalty=-100)
    want_final = int(tr.ends[bi[0]])
    lower, upper = O.boundaries(sizes)
    bad = want[0].copy()
    t = next(t for t in range(10, 90) if bad[t] == bad[t + 1])           # inside a run: break it with a jump into the
    w = int(np.searchsorted(upper, bad[t]))                             # middle of ANOTHER word (not via END -> START)
    bad[t + 1] = lower[(w + 3) % len(sizes)] + 1
    ok, rel = adjudicate(ems[0], tr, -100, bad, want_final, want[0], want_final)
    assert not ok and rel == float("inf")
    # even with a tolerance that would excuse anything
    ok, rel = adjudicate(ems[0], tr, -100, bad, want_final, want[0], want_final, rtol=1e9)
    assert not ok
    # a final state that is not a word END is outside the grammar too
    ok, _ = adjudicate(ems[0], tr, -100, want[0], int(lower[0]), want[0], want_final, rtol=1e9)
    assert not ok


def test_exact_tie_between_two_identical_words_is_excused(golden):
    """Two copies of one word model with identical emission columns: the reference prefers the lowest word index
    (hidden_markov_model.py:533-559, argmax), a decoder that walks the second copy found a path of EQUAL score."""
    logA = golden["train_logA_1"]
    S = len(logA)
    tr = O.loop_trellis([logA, logA])
    half = _scores(np.random.default_rng(3), 3, 60, S)
    ems = [np.concatenate((h, h), axis=1) for h in half]
    _, _, want = O.viterbi_batch(ems, tr, penalty=-100)
    assert all(p[:-1].max() < S for p in want)                           # the oracle stays in the first copy
    got = [p + S for p in want]
    v = compare_loop_decodes(ems, tr, -100, [S, S], ("A", "B"), got)
    assert (v.identical_paths, v.excused, len(v.failed)) == (0, 3, 0) and v.worst_excused_rel == 0.0
    # ... and a tie that is only NEAR (second copy worse by 1e-3 of the score) is not
    worse = [np.concatenate((h, h + np.float32(1e-3 * h.mean())), axis=1) for h in half]
    v = compare_loop_decodes(worse, tr, -100, [S, S], ("A", "B"), got)
    assert len(v.failed) == 3 and v.excused == 0


def test_true_states_restores_the_final_state(loop):
    """The reference's backtrace repeats s_{T-2} at T-1 (hidden_markov_model.py:574-580); the scored sequence ends in
    the best END state, which final_state_of_loop_path recovers from the word that holds s_{T-2}."""
    tr, sizes = loop
    ems = _scores(np.random.default_rng(4), 3, 80, sum(sizes))
    _, bi, want = O.viterbi_batch(ems, tr, penalty=-100)
    for p, b in zip(want, bi):
        assert p[-1] == p[-2]
        assert final_state_of_loop_path(p, sizes) == int(tr.ends[b])
        st = true_states(p, int(tr.ends[b]))
        assert st[-1] == int(tr.ends[b]) and np.array_equal(st[:-1], p[:-1])
