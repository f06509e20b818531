#!/usr/bin/env python
"""The reference's project-5 workflow on synthetic data, written against the package API exactly the
way the reference's drivers use it (scripts/project5_train_no_empty.py, project5_test_ndigits_with_sil.py):

  SignalSeparation.remove_empty_batch -> MFCC.batch -> HiddenMarkovModelTrainable.from_data(...).save
  -> HiddenMarkovModelInference.from_folder -> poke _log_transition_probability_between_words
  -> concurrent.futures.ProcessPoolExecutor().map(partial(_make_prediction, hmm), labeled.items())

Every hot-path call lands in the CUDA kernels; the process pool works because importing the package
switches multiprocessing to "forkserver" (torch preloaded, no CUDA in the server) when the engine is created.  Run on a GPU box:  python examples/digits_pipeline.py [workdir]
"""
import concurrent.futures
import functools
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cs-304-speech-recognition-code_b200"))

from loe_speech_recognition import (MFCC, TI_DIGITS_LABELS, HiddenMarkovModelInference, HiddenMarkovModelTrainable,  # noqa: E402
                                    SignalSeparation)
from loe_speech_recognition.synthetic import DIGITS, synth_string  # noqa: E402


def _make_prediction(hmm_inference, label_and_signals):
    label, signals = label_and_signals
    return [label] * len(signals), [hmm_inference.predict(s) for s in signals]


def main():
    work = sys.argv[1] if len(sys.argv) > 1 else tempfile.mkdtemp(prefix="loe_b200_")
    model_dir = os.path.join(work, "big_model_speech_only")
    rng = np.random.default_rng(0)
    # "recordings": one digit with leading / trailing silence, like a TIDIGITS isolated-digit file
    train = {w: [synth_string(rng, [w]) for _ in range(12)] for w in DIGITS}
    signal_separation = SignalSeparation(sample_rate=16000, speech_high_threshold=0.06, speech_low_threshold=0.01)
    for label in TI_DIGITS_LABELS:
        speech_only = signal_separation.remove_empty_batch(train[label])
        mfccs = MFCC.batch(speech_only, sample_rate=16000)
        hmm = HiddenMarkovModelTrainable.from_data(label, mfccs, num_of_states=5, max_iterations=10,
                                                   isMultiProcessingTraining=True, isTqdm=False)
        hmm.save(model_dir)
    hmm = HiddenMarkovModelTrainable.from_data("S", MFCC.batch(signal_separation.get_all_noises(), sample_rate=16000),
                                               num_of_states=3, max_iterations=10, isTqdm=False)
    hmm.save(model_dir)

    models_to_load = list(TI_DIGITS_LABELS.keys()) + ["S"]
    hmm_inference = HiddenMarkovModelInference.from_folder(model_dir, models_to_load)
    hmm_inference._log_transition_probability_between_words = -100
    test = {}
    for _ in range(12):
        digits = [DIGITS[int(i)] for i in rng.integers(0, 11, size=3)]
        test.setdefault("".join(digits), []).append(synth_string(rng, digits))
    labeled = {label: MFCC.batch(signals, sample_rate=16000) for label, signals in test.items()}
    truth, pred = [], []
    with concurrent.futures.ProcessPoolExecutor(max_workers=2) as executor:
        for t, p in executor.map(functools.partial(_make_prediction, hmm_inference), labeled.items()):
            truth.extend(t)
            pred.extend(p)
    batch_pred = []
    for label, feats in labeled.items():
        batch_pred.extend(hmm_inference.predict_batch(feats))
    assert batch_pred == pred, "process-pool predictions differ from the batched entry point"
    acc = sum(a == b for a, b in zip(truth, pred)) / len(truth)
    print(f"decoded {len(truth)} strings through a process pool, exact-string accuracy {acc:.2f}; models in {model_dir}")
    return acc


if __name__ == "__main__":
    main()
