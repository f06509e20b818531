#!/usr/bin/env python
"""Supplementary benchmark: segmental K-means training (BASELINE.json configs[2]).

    python bench_train.py [--utts-per-word 9091] [--iters 5]
    torchrun --nproc-per-node N bench_train.py ...        (utterances sharded by rank, one all-reduce / iteration)

11 digit HMMs x 5 states, isolated utterances (0.30-0.45 s, T ~ 30-46); a pool of distinct synthetic
utterances per word is pushed through the MFCC kernel once and tiled on the device to the requested
size.  One "iteration" = emission + Viterbi alignment + align/statistics kernels (+ all-reduce) + the
host M-step (scipy rebuild of 5 Gaussians), for one word model.  Prints one JSON line (rank 0).
This is NOT the driver's bench contract (that is bench.py / decode); it documents the training path.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "cs-304-speech-recognition-code_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts-per-word", type=int, default=9091)
    ap.add_argument("--pool", type=int, default=64)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--mode", default="both", choices=["per-word", "batch", "both"])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from loe_speech_recognition import HiddenMarkovModelTrainable
    from loe_speech_recognition._engine import Batch, get_engine
    from loe_speech_recognition.synthetic import DIGITS, isolated_corpus
    eng = get_engine()
    corpus = isolated_corpus(seed=3, n_per_word=args.pool, words=DIGITS)
    per_rank = (args.utts_per_word + world - 1) // world
    total_frames = 0
    t_iter, t_dev = [], []
    means_sum = 0.0
    for w in (DIGITS if args.mode != "batch" else ()):
        b = eng.mfcc(corpus[w])
        first = b.feat[: int(b.frm_off_host[1])].cpu().numpy()
        # tile the pool on the device to this rank's shard
        reps = (per_rank + args.pool - 1) // args.pool
        lens = np.tile(np.diff(b.frm_off_host), reps)[:per_rank]
        feat = b.feat.repeat(reps, 1)[: int(lens.sum())].contiguous()
        off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
        batch = Batch(feat, eng._to_dev(off), off, per_rank, int(lens.max()))
        total_frames += int(off[-1]) * world
        m = HiddenMarkovModelTrainable(w, isTqdm=False)
        m._means, m._covariances, m._transition_probs = m._init_parameters(first, 5)
        m._update_inference_weights()
        for it in range(args.iters):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            try:
                stats, counts = m._device_statistics(batch)
                t1 = time.perf_counter()
                m._update_from_statistics(stats, counts, shift=m._means.astype(np.float64))
            except HiddenMarkovModelTrainable.HMMTrainConverge:
                t1 = time.perf_counter()
            m._update_inference_weights()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            if it > 0:
                t_iter.append(t2 - t0); t_dev.append(t1 - t0)
        means_sum += float(np.abs(m._means).sum())
    batch_line = None
    if args.mode in ("batch", "both"):
        # all 11 words in one device pass per iteration (HiddenMarkovModelTrainable.from_data_batch)
        feats_by_word = {}
        for w in DIGITS:
            b = eng.mfcc(corpus[w])
            flat = b.feat.cpu().numpy()
            pool = [flat[b.frm_off_host[i]:b.frm_off_host[i + 1]] for i in range(len(corpus[w]))]
            reps = (args.utts_per_word + args.pool - 1) // args.pool
            feats_by_word[w] = (pool * reps)[:args.utts_per_word]          # from_data_batch shards by rank itself
        HiddenMarkovModelTrainable.from_data_batch(feats_by_word, num_of_states=5, max_iterations=2)    # warm-up (lazy kernel loading)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        models = HiddenMarkovModelTrainable.from_data_batch(feats_by_word, num_of_states=5, max_iterations=args.iters)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        models = HiddenMarkovModelTrainable.from_data_batch(feats_by_word, num_of_states=5, max_iterations=1)
        torch.cuda.synchronize()
        d1 = time.perf_counter() - t1
        per_iter = (dt - d1) / max(1, args.iters - 1)       # upload + setup cancel out
        chk = float(sum(np.abs(m._means).sum() for m in models.values()))
        batch_line = {"ms_per_iteration_all_words": per_iter * 1e3, "ms_setup_and_first_iteration": d1 * 1e3,
                      "utterances_per_s_per_iteration": args.utts_per_word * len(DIGITS) / per_iter, "means_checksum": chk}
    if args.mode == "batch":
        t_iter, t_dev = [float("nan")], [float("nan")]
    if world > 1:
        t = torch.tensor([np.mean(t_iter), np.mean(t_dev), means_sum], dtype=torch.float64, device=eng.device)
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mn = t.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        it_s, dev_s = float(mx[0]), float(mx[1])
        identical = bool(mx[2] == mn[2])
    else:
        it_s, dev_s, identical = float(np.mean(t_iter)), float(np.mean(t_dev)), True
    if rank == 0:
        utts = args.utts_per_word * len(DIGITS)
        print(json.dumps({
            "metric": "segmental K-means iteration (one word model)", "n_gpus": world,
            "utterances_total": utts, "frames_total": total_frames, "utts_per_word": args.utts_per_word,
            "ms_per_iteration_per_word": it_s * 1e3, "ms_device_part": dev_s * 1e3,
            "utterances_per_s_per_iteration": args.utts_per_word / it_s,
            "frames_per_s": total_frames / len(DIGITS) / it_s,
            "models_identical_across_ranks": identical, "means_checksum": means_sum,
            "allreduce_payload_bytes": 8 * (5 * 820 + 25), "batched_trainer": batch_line,
            "data": "synthetic (pool of %d per word, tiled)" % args.pool}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
