import sys, numpy as np, time, threading, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "cs-304-speech-recognition-code_b200"))
from loe_speech_recognition import _native
lib = _native.load()
rng = np.random.default_rng(0)
a = rng.integers(-32768, 32768, 100_000_000).astype(np.float32); out = np.zeros(a.size, np.int16)
print("cpus", os.cpu_count())
for W in (1, 2, 4, 8, 16, 32):
    n = a.size
    def work(w):
        i0 = n * w // W; i1 = n * (w + 1) // W
        lib.loe_pcm_narrow_host(a.ctypes.data + 4 * i0, out.ctypes.data + 2 * i0, i1 - i0)
    best = 0
    for rep in range(3):
        t = time.perf_counter()
        th = [threading.Thread(target=work, args=(w,)) for w in range(W)]
        [x.start() for x in th]; [x.join() for x in th]
        best = max(best, a.nbytes / (time.perf_counter() - t) / 1e9)
    print(W, "threads GB/s (float bytes):", round(best, 2))
