"""Time loe_kmeans_dev (counting sort + accum2 + reduce2) on the configs[2] shape with the library given as argv[1]:
100 000 single-digit utterances of ~38 frames, 11 words x 5 states, frames credited in left-to-right runs.
python scratch/time_kmeans.py <lib.so> [reps]  -> ms per call, checksum, max error against float64 NumPy on two buckets."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cs-304-speech-recognition-code_b200"))
import ctypes, hashlib
from loe_speech_recognition import _native
libs = [a for a in sys.argv[1:] if a.endswith(".so")] or [_native.LIB_PATH]
reps = next((int(a) for a in sys.argv[1:] if a.isdigit()), 10)


def load(path):
    lib = ctypes.CDLL(os.path.abspath(path))
    for name in ("loe_kmeans_dev", "loe_kmeans_ws_doubles", "loe_last_error"):
        res, args = _native.SIGNATURES[name]
        getattr(lib, name).restype, getattr(lib, name).argtypes = res, args
    return lib


lib = load(libs[0])
dev = torch.device("cuda", 0)
rng = np.random.default_rng(5)
n, W, S, D = 100000, 11, 5, 39
lens = rng.integers(25, 52, size=n)
F = int(lens.sum())
word = rng.integers(0, W, size=n)
# left-to-right alignment: S runs of random lengths per utterance
cuts = np.sort(rng.integers(1, lens[:, None], size=(n, S - 1)), axis=1)
bucket = np.empty(F, dtype=np.uint16)
off = np.concatenate(([0], np.cumsum(lens)))
state = np.zeros(F, dtype=np.int64)
pos = np.arange(F) - np.repeat(off[:-1], lens)
for k in range(S - 1):
    state += pos >= np.repeat(cuts[:, k], lens)
bucket[:] = np.repeat(word, lens) * S + state
bucket[rng.random(F) < 0.01] = 0xFFFF                       # not credited
x = rng.normal(0, 2, size=(F, D)).astype(np.float32)
shift = rng.normal(size=(W * S, D)).astype(np.float32)
G = W * S
stride = 1 + D + D * (D + 1) // 2
xd = torch.from_numpy(x).to(dev)
bd = torch.from_numpy(bucket.view(np.int16)).to(dev)
sd = torch.from_numpy(shift).to(dev)
ws = torch.empty((int(lib.loe_kmeans_ws_doubles(F, G, D)),), dtype=torch.float64, device=dev)
stats = torch.empty((G, stride), dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream


iu = np.triu_indices(D)
refs = {}
for g in (0, G - 1):
    d = x[bucket == g].astype(np.float64) - shift[g]
    refs[g] = np.concatenate(([len(d)], d.sum(0), (d.T @ d)[iu]))
for path in libs:
    lib = load(path)

    def call():
        rc = lib.loe_kmeans_dev(xd.data_ptr(), bd.data_ptr(), F, D, G, sd.data_ptr(), ws.data_ptr(), stats.data_ptr(), stream)
        assert rc == 0, lib.loe_last_error()

    stats.zero_()
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = stats.cpu().numpy()
    err = max(float(np.max(np.abs(st[g] - ref) / (np.abs(ref) + 1e-8))) for g, ref in refs.items())
    print(f"{os.path.basename(path)}: {ms:.4f} ms per call ({F} frames), sha1 {hashlib.sha1(st.tobytes()).hexdigest()[:12]}, "
          f"max rel err {err:.2e}", flush=True)
