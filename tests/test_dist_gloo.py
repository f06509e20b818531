"""CPU: the N > 1 path with world_size 2 over gloo -- sharding and the one-collective-per-iteration
statistics exchange of segmental K-means, followed by the identical M-step on every rank."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "cs-304-speech-recognition-code_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from loe_speech_recognition import _dist, HiddenMarkovModelTrainable
    from test_host_logic import _numpy_stats
    from helpers import oracle_flat
    from oracle import hmm as O
    golden = np.load(os.path.join(ROOT, "tests", "golden", "golden_hmm.npz"))
    w = "4"
    feats = [golden[f"train_feat_{w}_{i}"] for i in range(8)]
    means, Us, lps, logA = oracle_flat(golden, w)
    tr = O.word_trellis(logA)
    paths = [O.viterbi(O.emission_scores(x, means, Us, lps), tr)[2] for x in feats]
    assert _dist.world() == (rank, world)
    mine = _dist.shard(list(range(8)))
    assert mine == list(range(8))[rank::world]
    shift = np.zeros((5, 39))
    stats, counts = _numpy_stats([feats[i] for i in mine], [paths[i] for i in mine], 5, shift)
    s, c = _dist.allreduce_stats(torch.from_numpy(stats), torch.from_numpy(counts.astype(np.int32)))
    m = HiddenMarkovModelTrainable(w)
    m._means = np.zeros((5, 39), np.float32)
    m._covariances = m._init_covariance(39, 5)
    m._update_from_statistics(s.numpy(), c.numpy().astype(np.int64), shift=shift)
    ref = O.mstep(feats, paths, 5)
    ok = (np.allclose(m._means, ref["means"], rtol=1e-6, atol=1e-6) and np.allclose(m._covariances, ref["covs"], rtol=1e-5, atol=1e-7)
          and np.array_equal(m._transition_probs.to_dense(), ref["trans"]) and np.array_equal(c.numpy(), ref["counts"]))
    # every rank must hold the same model bit for bit
    gathered = [torch.zeros(5, 39) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(m._means))
    ok = ok and all(torch.equal(g, gathered[0]) for g in gathered)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_stats_allreduce_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_shard_by_frames_balances():
    sys.path.insert(0, os.path.join(ROOT, "cs-304-speech-recognition-code_b200"))
    from loe_speech_recognition import _dist
    rng = np.random.default_rng(0)
    lens = rng.integers(290, 460, size=1000).tolist()
    parts = _dist.shard_by_frames(lens, 8)
    assert sorted(i for p in parts for i in p) == list(range(1000))
    loads = [sum(lens[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 460
    assert _dist.world() == (0, 1) and _dist.shard([1, 2, 3]) == [1, 2, 3]
