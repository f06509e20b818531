"""The diagonal-GMM scoring kernel alone (BASELINE.json configs[4] shape: 120 states x 16 mixtures) on synthetic features:
the command the `ncu --set full` capture of emission_gmm_tc_kernel under profiles/ is taken from.

    ncu --set full --clock-control none --import-source on -k regex:emission_gmm_tc -s 1 -c 1 -o gpurun_out/prof_gmm python profiles/gmm_for_ncu.py
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cs-304-speech-recognition-code_b200"))
from loe_speech_recognition._engine import get_engine  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
eng = get_engine()
rng = np.random.default_rng(4)
S, M, D = 120, 16, 39
means = rng.normal(0, 1.0, size=(S, 1, D)) + rng.normal(0, 0.4, size=(S, M, D))
variances = rng.uniform(0.3, 0.9, size=(S, M, D)) ** 2
w = rng.uniform(0.5, 1.0, size=(S, M)); w /= w.sum(1, keepdims=True)
gp = eng.pack_gmm(w, means, variances)
x = torch.randn((F, D), device=eng.device, dtype=torch.float32)
out = torch.empty((F, S), dtype=torch.float32, device=eng.device)
for _ in range(2):
    eng.emission_gmm(x, gp, "tc", out=out)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    eng.emission_gmm(x, gp, "tc", out=out)
torch.cuda.synchronize()
print("frames", F, "ms", (time.perf_counter() - t0) / 3 * 1e3)
