"""Per-call latency of the reference's single-utterance API through the drop-in (the calls the reference's scripts make):
MFCC(signal).feature_vector, HiddenMarkovModelInference.predict, ModelCollection.predict, HiddenMarkovModel.predict."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-304-speech-recognition-code_b200")); sys.path.insert(0, ROOT)
from loe_speech_recognition import HiddenMarkovModel, HiddenMarkovModelInference, HiddenMarkovModelTrainable, MFCC, ModelCollection
from loe_speech_recognition.synthetic import string_corpus
from loe_speech_recognition.transition_probability import LogTransitionProbabilities

g = np.load(os.path.join(ROOT, "tests", "golden", "golden_hmm.npz"))
order = ("1", "2", "3", "4", "5", "6", "7", "8", "9", "O", "S", "Z")
models = []
for w in order:
    m = HiddenMarkovModel(w)
    m._multivariate_normals = HiddenMarkovModelTrainable.get_multivariate_normals(g[f"train_means_{w}"], g[f"train_covs_{w}"])
    m._log_transition_probs = LogTransitionProbabilities.from_dense(g[f"train_logA_{w}"])
    models.append(m)
inf = HiddenMarkovModelInference.from_models(models)
inf._log_transition_probability_between_words = -100
mc = ModelCollection()
mc._models = [m for m in models if m.label != "S"] if hasattr(models[0], "label") else models[:11]
utts, _ = string_corpus(seed=5, n_utts=8, n_digits=7)
iso, _ = string_corpus(seed=6, n_utts=8, n_digits=1)

def bench(name, fn, n=200):
    for _ in range(10): fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    dt = (time.perf_counter() - t0) / n
    print(f"{name:55s} {dt * 1e6:9.1f} us / call")
    return dt

feats = MFCC.batch(utts, 16000); fiso = MFCC.batch(iso, 16000)
print("frames:", feats[0].shape, fiso[0].shape)
bench("MFCC(signal 7 digits).feature_vector", lambda: MFCC(utts[0], 16000).feature_vector)
bench("MFCC.batch([1 signal])", lambda: MFCC.batch(utts[:1], 16000))
bench("HiddenMarkovModelInference.predict (58 states)", lambda: inf.predict(feats[0]))
bench("ModelCollection.predict (11 x 5 states)", lambda: mc.predict(fiso[0]))
bench("HiddenMarkovModel.predict (5 states)", lambda: models[0].predict(fiso[0]))
bench("predict_batch(8 utterances)", lambda: inf.predict_batch(feats))
if "--profile" in sys.argv:
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(200): inf.predict(feats[0])
    pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
