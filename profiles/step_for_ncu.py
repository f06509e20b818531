"""The device-resident decode step of bench.py (BASELINE.json configs[1], 10 000 utterances) and nothing else:
the command the `ncu --set full` captures under profiles/ are taken from.

    ncu --set full --clock-control none --import-source on -k regex:"mfcc|emission|viterbi" -s 5 -c 4 \
        -o gpurun_out/prof python profiles/step_for_ncu.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cs-304-speech-recognition-code_b200"))
import bench  # noqa: E402
from loe_speech_recognition import HiddenMarkovModel, HiddenMarkovModelInference, HiddenMarkovModelTrainable  # noqa: E402
from loe_speech_recognition._engine import get_engine  # noqa: E402
from loe_speech_recognition.transition_probability import LogTransitionProbabilities  # noqa: E402

n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
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
utts, _ = bench.make_corpus(100, n_utts, 500)
lens = np.array([len(u) for u in utts], dtype=np.int64)
pcm_off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
frames = 1 + lens // 160
frm_off = np.concatenate(([0], np.cumsum(frames))).astype(np.int64)
F, n = int(frm_off[-1]), len(utts)
pcm = torch.from_numpy(np.concatenate(utts).astype(np.float32)).to(eng.device)
po, fo = eng._to_dev(pcm_off), eng._to_dev(frm_off)
gp, tp = inf._packs()
skip = inf._model_boundaries._labels.index("S")
image = eng.image_buffers(F)           # the decode hand-off of round 2: the cepstrum kernel writes the emission kernel's A operand
for _ in range(2):
    eng.mfcc_device(pcm, po, fo, n, F, int(frames.max()), int(frames.min()), 16000, image=image, want_feat=False)
    scores = eng.emission_image(image, F, gp)
    eng.viterbi(scores, fo, n, int(frames.max()), F, tp, loop=True, penalty=float(bench.PENALTY), penalty_f64=False,
                want_end_scores=False, labels=(skip, 32))
torch.cuda.synchronize()
print("frames", F)
