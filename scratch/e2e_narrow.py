"""e2e of loe_decoder_decode_host with float32 PCM: narrowing off vs forced on (bench workload)."""
import os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(__file__), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cs-304-speech-recognition-code_b200"))
import bench
from loe_speech_recognition import HiddenMarkovModel, HiddenMarkovModelInference, HiddenMarkovModelTrainable
from loe_speech_recognition._decoder import PinnedBuffer
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
utts, _ = bench.make_corpus(100, 10000, 500)
off = np.concatenate(([0], np.cumsum([len(u) for u in utts]))).astype(np.int64)
pin = PinnedBuffer(int(off[-1]), np.float32)
pin.array[:] = np.concatenate(utts)
ref = None
for threads, min_gbps in (("0", None), ("16", "0"), ("8", "0"), ("16", None)):
    os.environ["LOE_B200_NARROW_THREADS"] = threads
    if min_gbps is None: os.environ.pop("LOE_B200_NARROW_MIN_GBPS", None)
    else: os.environ["LOE_B200_NARROW_MIN_GBPS"] = min_gbps
    inf.__dict__.pop("_native_decoder", None)
    for _ in range(3): s = inf.decode_pcm_host(pin.array, off)
    t = time.perf_counter()
    for _ in range(8): s = inf.decode_pcm_host(pin.array, off)
    dt = (time.perf_counter() - t) / 8
    if ref is None: ref = s
    print(f"threads={threads} min_gbps={min_gbps}: {dt*1e3:.1f} ms/step, {len(utts)/dt:.0f} utt/s, rate {inf.native_decoder().narrow_rate():.1f} GB/s, same strings {s == ref}")
