"""Small end-to-end decode (MFCC f32 + s16, ragged lengths, FP16 / TF32 emission, loop Viterbi, C decoder) for
compute-sanitizer memcheck / racecheck runs."""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(__file__), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cs-304-speech-recognition-code_b200"))
import bench
from loe_speech_recognition import HiddenMarkovModel, HiddenMarkovModelInference, HiddenMarkovModelTrainable, MFCC
from loe_speech_recognition._engine import get_engine
from loe_speech_recognition.transition_probability import LogTransitionProbabilities
params = bench.golden_params()
models = []
for w in bench.LOOP_ORDER:
    m = HiddenMarkovModel(w)
    m._multivariate_normals = HiddenMarkovModelTrainable.get_multivariate_normals(params[w][0], params[w][1])
    m._log_transition_probs = LogTransitionProbabilities.from_dense(params[w][2])
    models.append(m)
inf = HiddenMarkovModelInference.from_models(models)
inf._log_transition_probability_between_words = bench.PENALTY
utts, _ = bench.make_corpus(7, 12, 12)
utts = [u[: len(u) - i * 37] for i, u in enumerate(utts)] + [utts[0][:1441], utts[1][:1599]]
feats = MFCC.batch(utts, 16000)
feats16 = MFCC.batch([np.round(u).astype(np.int16) for u in utts], 16000)
eng = get_engine()
gp, tp = inf._packs()
for prec in ("h16", "tc", "fp32"):
    eng.emission(eng._to_dev(np.concatenate(feats)), gp, prec)
s1 = inf.decode_pcm_batch(utts)
off = np.concatenate(([0], np.cumsum([len(u) for u in utts]))).astype(np.int64)
s2 = inf.decode_pcm_host(np.concatenate(utts).astype(np.float32), off, n_chunks=3)
eng.torch.cuda.synchronize()
print("ok", s1 == s2, len(s1))
