#!/usr/bin/env python
"""Supplementary benchmark: MFCC front-end sweep (BASELINE.json configs[3]).

    python bench_mfcc.py [--utts 1000000] [--chunk 100000] [--pcm f32|s16] [--config reference|spec]

Synthetic utterances with lengths uniform in [16 000, 64 000] samples (1-4 s at 16 kHz), generated ON
THE DEVICE (the full float32 corpus of 1 M utterances would be 160 GB); processed in chunks of
`--chunk` utterances whose PCM (and features) stay resident in HBM.  Reports utterances/s, frames/s
and achieved HBM GB/s of the two MFCC kernels (algorithmic bytes: 4 or 2 B/sample read + 156 B/frame
written), for the reference parameter set (n_fft 320, hop 160, Hann, 40 mel, 13 ceps + delta + delta-delta) or the
"spec" set of configs[3] (512-point FFT, 400-sample Hamming, pre-emphasis 0.97, ln, CMN).
bench.py runs a bounded sample of both sets and reports them under "mfcc_sweep" (this script runs the full 1 M).
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


def run_sweep(eng, utts=1000000, chunk=100000, pcm_kind="f32", config_name="reference", warm=3):
    """One pass over `utts` synthetic utterances (generated on the device chunk by chunk); returns the result dict.
    config_name: "reference" (mfcc.py:31-34: n_fft 320, Hann, dB, per-frame normalisation; loe_mfcc_dev) or "spec"
    (BASELINE configs[3]: 512-point FFT, 400-sample Hamming, pre-emphasis 0.97, ln, CMN; loe_mfcc_ex_dev)."""
    import torch
    from loe_speech_recognition import MFCCConfig
    dev = eng.device
    cfg = MFCCConfig.spec() if config_name == "spec" else MFCCConfig()
    rng = np.random.default_rng(0)
    n_chunks = (utts + chunk - 1) // chunk
    total_ms, total_frames, total_samples, done = 0.0, 0, 0, 0
    peak_bytes = 0
    for c in range(n_chunks):
        n = min(chunk, utts - done)
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
        if pcm_kind == "s16":
            pcm = pcm.to(torch.int16)
        feat = torch.empty((F, 39), dtype=torch.float32, device=dev)
        mel_ws = torch.empty((F, 40), dtype=torch.float32, device=dev)
        po, fo = eng._to_dev(pcm_off), eng._to_dev(frm_off)
        if cfg.is_reference:
            utt_max = torch.empty((n,), dtype=torch.float32, device=dev)
            run = lambda: eng.mfcc_device(pcm, po, fo, n, F, int(frames.max()), int(frames.min()), 16000, out=feat, mel_ws=mel_ws, utt_max=utt_max)
        else:
            ceps_ws = torch.empty((F, 13), dtype=torch.float32, device=dev)
            utt_stat = torch.empty((n, 26), dtype=torch.float32, device=dev)
            run = lambda: eng.mfcc_ex_device(pcm, po, fo, n, F, int(frames.max()), int(frames.min()), 16000, cfg, out=feat,
                                             mel_ws=mel_ws, ceps_ws=ceps_ws, utt_stat=utt_stat)
        if c == 0:
            for _ in range(warm):
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
    bps = 4 if pcm_kind == "f32" else 2
    alg = bps * total_samples + 156 * total_frames
    return {"metric": "MFCC front-end sweep", "parameter_set": config_name, "utterances": done, "frames": total_frames,
            "samples": total_samples, "pcm": pcm_kind, "chunk_utterances": chunk, "ms_total": total_ms,
            "utterances_per_s": done / (total_ms * 1e-3), "frames_per_s": total_frames / (total_ms * 1e-3),
            "algorithmic_GBps": alg / (total_ms * 1e-3) / 1e9, "algorithmic_bytes": alg,
            "peak_hbm_allocated_GB": peak_bytes / 1e9, "data": "synthetic 1-4 s utterances (generated on the device)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=1000000)
    ap.add_argument("--chunk", type=int, default=100000)
    ap.add_argument("--pcm", default="f32", choices=["f32", "s16"])
    ap.add_argument("--config", default="reference", choices=["reference", "spec"])
    args = ap.parse_args()
    from loe_speech_recognition._engine import get_engine
    print(json.dumps(run_sweep(get_engine(), args.utts, args.chunk, args.pcm, args.config)))


if __name__ == "__main__":
    main()
