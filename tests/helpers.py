"""Shared test helpers: rebuild reference-equivalent model objects from the golden arrays."""
import numpy as np

WORDS = ("1", "2", "3", "4", "5", "6", "7", "8", "9", "O", "Z", "S")
LOOP_ORDER = ("1", "2", "3", "4", "5", "6", "7", "8", "9", "O", "S", "Z")
N_STATES = {**{w: 5 for w in WORDS}, "S": 3}


def trained_word_model(golden, w):
    """HiddenMarkovModel equal to the reference-trained model of word w (means/covs/logA from the
    golden file; scipy rebuilds the frozen Gaussians exactly as the reference did)."""
    from loe_speech_recognition.hidden_markov_model import HiddenMarkovModel, HiddenMarkovModelTrainable
    from loe_speech_recognition.transition_probability import LogTransitionProbabilities
    m = HiddenMarkovModel(w)
    m._multivariate_normals = HiddenMarkovModelTrainable.get_multivariate_normals(golden[f"train_means_{w}"], golden[f"train_covs_{w}"])
    m._log_transition_probs = LogTransitionProbabilities.from_dense(golden[f"train_logA_{w}"])
    return m


def oracle_flat(golden, w):
    from oracle import hmm as O
    means = golden[f"train_means_{w}"]; covs = golden[f"train_covs_{w}"]
    packs = [O.gaussian_pack(means[s], covs[s]) for s in range(len(means))]
    return (np.array([p[0] for p in packs]), np.array([p[1] for p in packs]), np.array([p[2] for p in packs]),
            golden[f"train_logA_{w}"])


def rel_close(a, b, rtol=1e-4, atol=0.0):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    return bool(np.all(both_inf | (np.abs(a - b) <= rtol * np.abs(b) + atol)))


def emission_close(got, ref, log_pdets, dim=39, rtol=1e-4):
    """Parity bar for log-likelihoods: 1e-4 relative, where the scale of a score
    cst_s - maha/2 is max(|score|, |cst_s|): the two terms cancel (scores cross zero), so a bound
    relative to the difference alone is not meaningful in float32.  The float64 path is held to
    1e-6 of |score| itself (see test_emission_matches_scipy)."""
    cst = np.abs(-0.5 * (dim * np.log(2 * np.pi) + np.asarray(log_pdets, dtype=np.float64)))[None, :]
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    return bool(np.all(np.abs(got - ref) <= rtol * np.maximum(np.abs(ref), cst)))
