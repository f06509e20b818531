"""CPU: pin the oracle files against (a) each other everywhere and (b) the UNMODIFIED reference
when /root/reference exists (authoring container; skipped on the GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import LOOP_ORDER, oracle_flat
from oracle import hmm as O
from oracle import ref_port
from oracle.ref_import import reference_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _port_model(golden, penalty):
    return ref_port.PortModel([golden[f"train_means_{w}"] for w in LOOP_ORDER], [golden[f"train_covs_{w}"] for w in LOOP_ORDER],
                              [golden[f"train_logA_{w}"] for w in LOOP_ORDER], list(LOOP_ORDER), penalty)


@pytest.mark.parametrize("penalty", [-100, np.log(0.005)])
def test_port_matches_vectorised_oracle_and_golden(golden, penalty):
    model = _port_model(golden, penalty)
    name = "int" if penalty == -100 else "f64"
    x = golden["loop_feat_6"][:60]                       # 60 frames x 58 states x scipy call: ~0.2 s
    flat = [oracle_flat(golden, w) for w in LOOP_ORDER]
    em = O.emission_scores(x, np.concatenate([f[0] for f in flat]), np.concatenate([f[1] for f in flat]),
                           np.concatenate([f[2] for f in flat]))
    es, bi, path = O.viterbi(em, O.loop_trellis([f[3] for f in flat]), penalty=penalty)
    score, ppath = ref_port.loop_viterbi(model, x)
    assert score == es[bi] and np.array_equal(ppath, path)
    sizes = [f[3].shape[0] for f in flat]
    assert ref_port.labels_from_path(model, ppath) == "".join(O.get_labels(path, sizes, list(LOOP_ORDER)))
    # full-length golden path -> same string as the reference printed
    assert ref_port.labels_from_path(model, golden[f"loop_path_{name}_0"]) == str(golden[f"loop_strings_{name}"][0])


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
def test_port_and_oracle_match_real_reference(tmp_path):
    """Runs in a subprocess: the real package and the drop-in share the name loe_speech_recognition."""
    code = r'''
import sys, os, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
from oracle.ref_import import import_reference
from oracle import ref_port, hmm as O, mfcc as OM
R = import_reference()
assert R.__file__.startswith("/root/reference/")
g = np.load(os.path.join(%r, "tests", "golden", "golden_hmm.npz"))
order = ("1","2","3","4","5","6","7","8","9","O","S","Z")
from loe_speech_recognition.hidden_markov_model import HiddenMarkovModel, HiddenMarkovModelTrainable, HiddenMarkovModelInference
from loe_speech_recognition.transition_probability import LogTransitionProbabilities
from loe_speech_recognition.model_boundary import ModelBoundary
inf = HiddenMarkovModelInference()
ltp = LogTransitionProbabilities(); normals = []; mb = ModelBoundary()
for w in order:
    A = g["train_logA_" + w]
    one = LogTransitionProbabilities(A.shape[0])
    for i in range(A.shape[0]):
        for j in range(A.shape[0]):
            one[(i, j)] = A[i, j]
    ltp.append(one)
    normals.extend(HiddenMarkovModelTrainable.get_multivariate_normals(g["train_means_" + w], g["train_covs_" + w]))
    mb.append(A.shape[0])
mb.add_model_labels(list(order))
inf._log_transition_probs, inf._multivariate_normals, inf._model_boundaries = ltp, normals, mb
x = g["loop_feat_2"][:50]
for pen in (-100, np.log(0.005)):
    inf._log_transition_probability_between_words = pen
    score, path = inf._viterbi(x)
    pm = ref_port.PortModel([g["train_means_" + w] for w in order], [g["train_covs_" + w] for w in order],
                            [g["train_logA_" + w] for w in order], list(order), pen)
    ps, pp = ref_port.loop_viterbi(pm, x)
    assert ps == score and np.array_equal(pp, path), pen
    assert ref_port.labels_from_path(pm, pp) == inf.predict(x)
print("OK")
''' % (ROOT, ROOT, ROOT)
    out = subprocess.check_output([sys.executable, "-c", code], text=True, cwd=str(tmp_path))
    assert out.strip().endswith("OK")
