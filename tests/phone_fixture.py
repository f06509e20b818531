"""Synthetic phone-level task shaped like BASELINE.json configs[4]: a pronunciation lexicon for the eleven digits plus
silence, 3-state left-to-right phone HMMs, M-component diagonal GMMs per state, and feature sequences sampled from the
model (so that decoding is non-trivial and self-consistent)."""
import numpy as np

LEXICON = {
    "1": ["w", "ah", "n"], "2": ["t", "uw"], "3": ["th", "r", "iy"], "4": ["f", "ao", "r"], "5": ["f", "ay", "v"],
    "6": ["s", "ih", "k", "s"], "7": ["s", "eh", "v", "ah", "n"], "8": ["ey", "t"], "9": ["n", "ay", "n"],
    "O": ["ow"], "S": ["sil"], "Z": ["z", "iy", "r", "ow"],
}
ORDER = ["1", "2", "3", "4", "5", "6", "7", "8", "9", "O", "S", "Z"]       # sorted(os.listdir), like the reference's loop


def make_phone_task(seed=0, n_mix=16, n_utts=40, n_digits=4, dim=39):
    rng = np.random.default_rng(seed)
    phones = sorted({p for ps in LEXICON.values() for p in ps})
    phone_col = {p: 3 * i for i, p in enumerate(phones)}
    S = 3 * len(phones)
    centers = rng.normal(0, 3.0, size=(S, 1, dim))
    means = centers + rng.normal(0, 0.7, size=(S, n_mix, dim))
    variances = rng.uniform(0.2, 1.0, size=(S, n_mix, dim))
    w = rng.uniform(0.5, 1.0, size=(S, n_mix))
    weights = w / w.sum(axis=1, keepdims=True)
    phone_logA, phone_log_exit = {}, {}
    for p in phones:
        stay = rng.uniform(0.6, 0.8, size=3)
        a = np.zeros((3, 3))
        for i in range(3):
            a[i, i] = stay[i]
            if i < 2:
                a[i, i + 1] = 1 - stay[i]
        with np.errstate(divide="ignore"):
            phone_logA[p] = np.log(a).astype(np.float32)
        phone_log_exit[p] = float(np.log(1 - stay[2]))
    feats, truth = [], []
    for _ in range(n_utts):
        digits = [ORDER[i] for i in rng.choice([i for i in range(12) if ORDER[i] != "S"], size=n_digits)]
        seq = ["S"] + [x for d in digits for x in (d, "S")]
        frames = []
        for wlab in seq:
            for p in LEXICON[wlab]:
                for j in range(3):
                    s = phone_col[p] + j
                    for _ in range(int(rng.integers(3, 7))):
                        m = rng.choice(n_mix, p=weights[s])
                        frames.append(means[s, m] + rng.normal(size=dim) * np.sqrt(variances[s, m]))
        feats.append(np.asarray(frames, dtype=np.float32))
        truth.append("".join(digits))
    return dict(weights=weights, means=means, variances=variances, phone_logA=phone_logA, phone_log_exit=phone_log_exit,
                phone_col=phone_col, lexicon=LEXICON, order=ORDER, feats=feats, truth=truth, phones=phones)
