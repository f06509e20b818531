"""Does pipelining the device-resident decode step over K sub-batches on K streams fill the kernel tails?
python scratch/two_streams.py -> ms per 10 000 utterances: one stream / K sub-batches on one stream / K sub-batches on K streams
(the strings of the split runs are compared with the single-batch run)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cs-304-speech-recognition-code_b200"))
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
gp, tp = inf._packs()
skip = inf._model_boundaries._labels.index("S")


class Part:
    def __init__(self, utts):
        lens = np.array([len(u) for u in utts], dtype=np.int64)
        frames = 1 + lens // 160
        self.n = len(utts)
        self.F = int(frames.sum())
        self.max_t, self.min_t = int(frames.max()), int(frames.min())
        self.pcm = torch.from_numpy(np.concatenate(utts).astype(np.float32)).to(eng.device)
        self.po = eng._to_dev(np.concatenate(([0], np.cumsum(lens))).astype(np.int64))
        self.fo = eng._to_dev(np.concatenate(([0], np.cumsum(frames))).astype(np.int64))
        self.image = eng.image_buffers(self.F)
        self.mel = torch.empty((self.F, 40), dtype=torch.float32, device=eng.device)
        self.umax = torch.empty((self.n,), dtype=torch.float32, device=eng.device)
        self.scores = torch.empty((self.F, gp.n_states), dtype=torch.float32, device=eng.device)

    def run(self):
        eng.mfcc_device(self.pcm, self.po, self.fo, self.n, self.F, self.max_t, self.min_t, 16000, mel_ws=self.mel,
                        utt_max=self.umax, image=self.image, want_feat=False)
        eng.emission_image(self.image, self.F, gp, out=self.scores)
        return eng.viterbi(self.scores, self.fo, self.n, self.max_t, self.F, tp, loop=True, penalty=float(bench.PENALTY),
                           penalty_f64=False, want_end_scores=False, labels=(skip, 32))


def words_of(res):
    words, count = res[4].cpu().numpy(), res[5].cpu().numpy()
    return [bytes(words[i, :count[i]].astype(np.uint8)) for i in range(len(count))]


def timed(fn, reps=10):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


whole = Part(utts)
t1, res = timed(whole.run)
want = words_of(res)
print(f"1 batch, 1 stream: {t1:.3f} ms", flush=True)
main = torch.cuda.current_stream()
for K in (2, 3, 4):
    cut = [len(utts) * k // K for k in range(K + 1)]
    parts = [Part(utts[cut[k]:cut[k + 1]]) for k in range(K)]
    streams = [torch.cuda.Stream() for _ in range(K)]
    done = [torch.cuda.Event() for _ in range(K)]

    def serial():
        return [p.run() for p in parts]

    def fanned():
        start = torch.cuda.Event()
        start.record(main)
        outs = []
        for p, s, d in zip(parts, streams, done):
            s.wait_event(start)
            with torch.cuda.stream(s):
                outs.append(p.run())
                d.record(s)
        for d in done:
            main.wait_event(d)
        return outs

    ts, outs = timed(serial)
    assert sum((words_of(o) for o in outs), []) == want
    tf, outs = timed(fanned)
    torch.cuda.synchronize()
    assert sum((words_of(o) for o in outs), []) == want
    print(f"{K} sub-batches: one stream {ts:.3f} ms, {K} streams {tf:.3f} ms", flush=True)
