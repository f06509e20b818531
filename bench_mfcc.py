#!/usr/bin/env python
"""Supplementary benchmark: MFCC front-end sweep (BASELINE.json configs[3]).

    python bench_mfcc.py [--utts 1000000] [--chunk 100000] [--pcm f32|s16]

Synthetic utterances with lengths uniform in [16 000, 64 000] samples (1-4 s at 16 kHz), generated ON
THE DEVICE (the full float32 corpus of 1 M utterances would be 160 GB); processed in chunks of
`--chunk` utterances whose PCM (and features) stay resident in HBM.  Reports utterances/s, frames/s
and achieved HBM GB/s of the two MFCC kernels (algorithmic bytes: 4 or 2 B/sample read + 156 B/frame
written), reference parameter set (n_fft 320, hop 160, Hann, 40 mel, 13 ceps + delta + delta-delta).
Not the driver's bench contract (that is bench.py).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "cs-304-speech-recognition-code_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=1000000)
    ap.add_argument("--chunk", type=int, default=100000)
    ap.add_argument("--pcm", default="f32", choices=["f32", "s16"])
    args = ap.parse_args()
    import torch
    from loe_speech_recognition._engine import get_engine
    eng = get_engine()
    dev = eng.device
    rng = np.random.default_rng(0)
    n_chunks = (args.utts + args.chunk - 1) // args.chunk
    total_ms, total_frames, total_samples, done = 0.0, 0, 0, 0
    peak_bytes = 0
    for c in range(n_chunks):
        n = min(args.chunk, args.utts - done)
        lens = rng.integers(16000, 64001, size=n).astype(np.int64)
        pcm_off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
        frames = 1 + lens // 160
        frm_off = np.concatenate(([0], np.cumsum(frames))).astype(np.int64)
        S, F = int(pcm_off[-1]), int(frm_off[-1])
        g = torch.Generator(device=dev); g.manual_seed(c)
        t = torch.arange(S, device=dev, dtype=torch.float32)
        pcm = (3000.0 * torch.sin(t * (2 * np.pi * 440.0 / 16000.0)) + 2000.0 * torch.sin(t * (2 * np.pi * 1730.0 / 16000.0)))
        pcm += 30.0 * torch.randn(S, device=dev, generator=g)
        pcm = pcm.round_()
        del t
        if args.pcm == "s16":
            pcm = pcm.to(torch.int16)
        feat = torch.empty((F, 39), dtype=torch.float32, device=dev)
        mel_ws = torch.empty((F, 40), dtype=torch.float32, device=dev)
        utt_max = torch.empty((n,), dtype=torch.float32, device=dev)
        po, fo = eng._to_dev(pcm_off), eng._to_dev(frm_off)
        run = lambda: eng.mfcc_device(pcm, po, fo, n, F, int(frames.max()), int(frames.min()), 16000, out=feat, mel_ws=mel_ws, utt_max=utt_max)
        if c == 0:
            for _ in range(3):
                run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
        total_frames += F; total_samples += S; done += n
        peak_bytes = max(peak_bytes, torch.cuda.max_memory_allocated(dev))
        assert bool(torch.isfinite(feat[:: max(1, F // 1000)]).all())
        del pcm, feat, mel_ws
    bps = 4 if args.pcm == "f32" else 2
    alg = bps * total_samples + 156 * total_frames
    print(json.dumps({"metric": "MFCC front-end sweep", "utterances": done, "frames": total_frames, "samples": total_samples,
                      "pcm": args.pcm, "chunk_utterances": args.chunk, "ms_total": total_ms,
                      "utterances_per_s": done / (total_ms * 1e-3), "frames_per_s": total_frames / (total_ms * 1e-3),
                      "algorithmic_GBps": alg / (total_ms * 1e-3) / 1e9, "algorithmic_bytes": alg,
                      "peak_hbm_allocated_GB": peak_bytes / 1e9, "data": "synthetic (generated on the device)"}))


if __name__ == "__main__":
    main()
