"""Time the four kernels of the device-resident decode step (bench.py's stage_ms) with the library given as argv[1]
(scratch/libs/*.so), one process per library: python scratch/time_step.py scratch/libs/x.so"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cs-304-speech-recognition-code_b200"))
from loe_speech_recognition import _native
_native.LIB_PATH = os.path.abspath(sys.argv[1])
import bench
from loe_speech_recognition import HiddenMarkovModel, HiddenMarkovModelInference, HiddenMarkovModelTrainable
from loe_speech_recognition._engine import get_engine
from loe_speech_recognition.transition_probability import LogTransitionProbabilities
eng = get_engine()
params = bench.golden_params()
models = []
for w in bench.LOOP_ORDER:
    m = HiddenMarkovModel(w)
    m._multivariate_normals = HiddenMarkovModelTrainable.get_multivariate_normals(params[w][0], params[w][1])
    m._log_transition_probs = LogTransitionProbabilities.from_dense(params[w][2])
    models.append(m)
inf = HiddenMarkovModelInference.from_models(models)
inf._log_transition_probability_between_words = bench.PENALTY
utts, _ = bench.make_corpus(100, 10000, 500)
lens = np.array([len(u) for u in utts], dtype=np.int64)
pcm_off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
frames = 1 + lens // 160
frm_off = np.concatenate(([0], np.cumsum(frames))).astype(np.int64)
F, n = int(frm_off[-1]), len(utts)
pcm = torch.from_numpy(np.concatenate(utts).astype(np.float32)).to(eng.device)
po, fo = eng._to_dev(pcm_off), eng._to_dev(frm_off)
gp, tp = inf._packs()
skip = inf._model_boundaries._labels.index("S")
image = eng.image_buffers(F)
def timed(fn, reps=10):
    for _ in range(3): out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out
t_mfcc, _ = timed(lambda: eng.mfcc_device(pcm, po, fo, n, F, int(frames.max()), int(frames.min()), 16000, image=image, want_feat=False))
t_em, scores = timed(lambda: eng.emission_image(image, F, gp))
t_vit, res = timed(lambda: eng.viterbi(scores, fo, n, int(frames.max()), F, tp, loop=True, penalty=float(bench.PENALTY), penalty_f64=False,
                                       want_end_scores=False, labels=(skip, 32)))
print(os.path.basename(sys.argv[1]), f"mfcc {t_mfcc:.3f} emission {t_em:.3f} viterbi {t_vit:.3f} ms")
