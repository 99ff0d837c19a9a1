"""CPU tests of the multi-rank host logic:
  * the sharded-training protocol (tests/sharded_model.py, the executable model of
    train.cu + dist_train.py) reproduces single-string training exactly, for every way of
    cutting the corpus -- in-process over many splits, and across two real processes over
    torch.distributed's gloo backend;
  * record sharding for encode covers every record exactly once."""
import os
import threading

import numpy as np
import pytest

import sharded_model as M


class ThreadGather:
    """all_gather for `world` threads in lockstep."""

    def __init__(self, world):
        self.world = world
        self.slots = [None] * world
        self.bar = threading.Barrier(world)

    def bind(self, rank):
        def ag(obj):
            self.slots[rank] = obj
            self.bar.wait()
            out = list(self.slots)
            self.bar.wait()
            return out
        return ag


def run_threads(text, cuts, m):
    bounds = [0] + list(cuts) + [len(text)]
    world = len(bounds) - 1
    g = ThreadGather(world)
    res = [None] * world

    def work(r):
        res[r] = M.train_rank(list(text[bounds[r]:bounds[r + 1]]), m, r, world, g.bind(r))

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return res


def check_against_oracle(oracle, text, cuts, m):
    res = run_threads(text, cuts, m)
    ids, pairs, counts, ntied = oracle.train_pairs(bytes(text), m)
    want = [(int(a), int(b), int(c), int(t)) for (a, b), c, t in zip(pairs.tolist(), counts.tolist(), ntied.tolist())]
    for tok, merges in res:
        assert merges == want
    cat = [t for tok, _ in res for t in tok]
    assert cat == ids.tolist()


def test_sharded_model_matches_single_string(oracle):
    rng = np.random.default_rng(0)
    for trial in range(25):
        n = int(rng.integers(2, 120))
        text = rng.integers(97, 100, size=n).astype(np.uint8).tolist()
        world = int(rng.integers(1, 6))
        cuts = sorted(int(c) for c in rng.integers(0, n + 1, size=world - 1))  # empty shards allowed
        check_against_oracle(oracle, text, cuts, int(rng.integers(1, 30)))


def test_sharded_model_runs_across_shards(oracle):
    # (x,x) runs that span one, two and three shard boundaries, odd and even lengths
    for run in (2, 3, 4, 5, 9, 16, 17):
        text = [98] + [97] * run + [99, 97, 97]
        for c1 in range(0, len(text) + 1, 2):
            for c2 in range(c1, len(text) + 1, 3):
                check_against_oracle(oracle, text, [c1, c2], 6)
    check_against_oracle(oracle, [97] * 33, [11, 11, 20], 8)   # all-x corpus, an empty shard
    check_against_oracle(oracle, [97, 98] * 20, [1, 2, 3], 5)  # single-token shards


def test_sharded_model_ecg_corpus(oracle, small_corpus):
    x, pct = small_corpus
    sym = oracle.quantize(x[:2, :2, :600], pct["percentile_1"], pct["percentile_99"]).reshape(-1).tolist()
    n = len(sym)
    check_against_oracle(oracle, sym, [n // 3, 2 * n // 3], 60)


def _gloo_worker(rank, world, port, text, m, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ecgbyte.dist_train import split_contiguous
        lo, hi = split_contiguous(len(text), world)[rank]

        def ag(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

        tok, merges = M.train_rank(list(text[lo:hi]), m, rank, world, ag)
        q.put((rank, tok, merges))
    finally:
        dist.destroy_process_group()


def test_sharded_protocol_over_gloo_world2(oracle, small_corpus):
    import torch.multiprocessing as mp
    x, pct = small_corpus
    sym = oracle.quantize(x[:1, :3, :500], pct["percentile_1"], pct["percentile_99"]).reshape(-1).tolist()
    m = 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, sym, m, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids, pairs, counts, ntied = oracle.train_pairs(bytes(sym), m)
    want = [(int(a), int(b), int(c), int(t)) for (a, b), c, t in zip(pairs.tolist(), counts.tolist(), ntied.tolist())]
    assert out[0][2] == want and out[1][2] == want
    assert out[0][1] + out[1][1] == ids.tolist()


def test_record_sharding_covers_everything():
    from ecgbyte.dist_train import split_contiguous
    for n in (0, 1, 7, 100000, 1000003):
        for world in (1, 2, 4, 8):
            parts = split_contiguous(n, world)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_model_many_shards(oracle):
    """The resident tail of the single-device trainer runs this protocol with one shard per CTA (hundreds of
    small chunks, some of them emptied by the merges): many shards, long (x,x) runs across many of them."""
    rng = np.random.default_rng(7)
    for trial in range(6):
        n = int(rng.integers(150, 400))
        text = rng.integers(97, 100, size=n).astype(np.uint8).tolist()
        lo = int(rng.integers(0, n - 60))
        text[lo:lo + 57] = [97] * 57                      # a run of x over ~6 shards
        world = int(rng.integers(24, 40))
        cuts = sorted(int(c) for c in rng.integers(0, n + 1, size=world - 1))
        check_against_oracle(oracle, text, cuts, int(rng.integers(20, 60)))
    even = [i * 10 for i in range(1, 30)]                 # equal chunks, as at the switch to the resident tail
    check_against_oracle(oracle, ([97] * 7 + [98, 97, 97, 99]) * 28, even, 40)
