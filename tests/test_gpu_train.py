"""K3-K5 parity: CUDA trainer vs the CPU oracle -- merge list, counts, tie log and the
merged id stream, bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _train_gpu(text, m, **kw):
    from ecgbyte.api import Trainer
    data = text if isinstance(text, (bytes, bytearray)) else np.ascontiguousarray(text, np.uint8).tobytes()
    tr = Trainer(max(len(data), 1), m, **kw)
    tr.load(data)
    pairs, counts, ntied = tr.run(m)
    return tr.ids(), pairs, counts, ntied


def _compare(oracle, text, m, fast=True, **kw):
    ids, pairs, counts, ntied = _train_gpu(text, m, **kw)
    o_ids, o_pairs, o_counts, o_ntied = oracle.train_pairs(text, m, fast=fast)
    np.testing.assert_array_equal(pairs, o_pairs)
    np.testing.assert_array_equal(counts, o_counts)
    np.testing.assert_array_equal(ntied, o_ntied)
    np.testing.assert_array_equal(ids, o_ids)


def test_train_kats(oracle):
    import rust_bpe
    # merge([a,a,a],(a,a)) -> [X,a]; overlapping counts: 'aaa' has (a,a) twice
    ids, vocab, merges = rust_bpe.byte_pair_encoding("aaa", 1, 1)
    assert ids == [256, 97] and merges == [([97, 97], 256)] and vocab[256] == "aa" and len(vocab) == 257
    # early stop when no pair is left (lib.rs:88-90)
    ids, vocab, merges = rust_bpe.byte_pair_encoding("ab", 5, 1)
    assert ids == [256] and len(merges) == 1
    ids, vocab, merges = rust_bpe.byte_pair_encoding("", 3, 1)
    assert ids == [] and merges == [] and len(vocab) == 256
    ids, vocab, merges = rust_bpe.byte_pair_encoding("a", 3, 1)
    assert ids == [97] and merges == []
    # tie rule: smallest (left, right) among equal counts
    ids, vocab, merges = rust_bpe.byte_pair_encoding("abcd", 1, 1)
    assert merges == [([97, 98], 256)]
    assert vocab[200] == "<200>"  # lib.rs:50-56
    with pytest.raises(TypeError):
        rust_bpe.byte_pair_encoding(b"abc", 1, 1)


@pytest.mark.parametrize("n,m", [(2, 4), (17, 10), (4096, 50), (4097, 50), (8193, 64), (100000, 300)])
def test_train_random_text(oracle, n, m):
    rng = np.random.default_rng(n)
    text = rng.integers(97, 101, size=n).astype(np.uint8)
    _compare(oracle, text, m, fast=n > 20000)


def test_train_runs_and_tile_edges(oracle):
    """(x,x) merges with runs that cross thread, tile and odd/even boundaries."""
    rng = np.random.default_rng(11)
    parts = []
    for _ in range(300):
        parts.append(np.full(int(rng.integers(1, 700)), 105, np.uint8))
        parts.append(rng.integers(104, 108, size=int(rng.integers(1, 4))).astype(np.uint8))
    parts.append(np.full(9001, 106, np.uint8))  # a run longer than two tiles
    text = np.concatenate(parts)
    _compare(oracle, text, 40)
    _compare(oracle, np.full(12289, 97, np.uint8), 14)  # a^n: every step is an (x,x) merge


def test_train_ecg_corpus(oracle, small_corpus):
    x, pct = small_corpus
    sym = oracle.quantize(x, pct["percentile_1"], pct["percentile_99"]).reshape(-1)
    _compare(oracle, sym, 600)


def test_train_resident_tail_dense_steps(oracle):
    """6e6 symbols fit the CTAs' shared memory from the first step (chunks of three tiles): the dense early
    merges, (x,x) steps included, all run through the resident in-place pass."""
    from ecgbyte import synth
    x = np.stack([synth.record(11, k, 5000) for k in range(100)])
    pct = synth.BENCH_PERCENTILES
    sym = oracle.quantize(x, pct["percentile_1"], pct["percentile_99"]).reshape(-1)
    _compare(oracle, sym, 250)
    # a little over the limit: the first steps stream, then the switch happens mid-run
    x2 = np.stack([synth.record(12, k, 5000) for k in range(140)])
    sym2 = oracle.quantize(x2, pct["percentile_1"], pct["percentile_99"]).reshape(-1)
    _compare(oracle, sym2, 120)


def test_train_table_overflow_is_loud(oracle):
    from ecgbyte import EcgbError
    rng = np.random.default_rng(2)
    text = rng.integers(0, 256, size=200000).astype(np.uint8)  # ~65k distinct pairs
    with pytest.raises(EcgbError):
        _train_gpu(text, 4, table_log2=10)


def test_byte_pair_encoding_reference_types(oracle, small_corpus):
    import rust_bpe
    x, pct = small_corpus
    text = oracle.quantize(x[:2], pct["percentile_1"], pct["percentile_99"]).tobytes().decode()
    ids, vocab, merges = rust_bpe.byte_pair_encoding(text, 120, 4)
    o = oracle.byte_pair_encoding(text, 120, fast=True)
    assert isinstance(ids, list) and isinstance(vocab, dict) and isinstance(merges, list)
    assert isinstance(merges[0], tuple) and isinstance(merges[0][0], list) and isinstance(merges[0][1], int)
    assert (ids, vocab, merges) == o
    assert len(vocab) == 256 + len(merges)


def test_train_config1_scale_matches_fixture():
    """BASELINE.json config 1: 1,000 PTB-XL-shaped records (6e7 symbols), 5,000 merges.
    The expected merge list / counts / tie log / final stream come from the oracle
    (tests/golden/ptbxl_1000_m5000.npz, oracle/make_table_fixture.py)."""
    import os
    from ecgbyte import synth
    from ecgbyte.api import Quantizer, Trainer
    f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ptbxl_1000_m5000.npz"))
    x = synth.corpus(0, 1000, 5000, np.float32)
    pct = synth.percentiles(x, seed=0)
    assert [pct["percentile_1"], pct["percentile_99"]] == f["pct"].tolist()
    q = Quantizer(pct, dtype=torch.float32)
    sym = q.quantize(torch.from_numpy(x).cuda()).reshape(-1)
    tr = Trainer(sym.numel(), 5000)
    tr.load(sym)
    pairs, counts, ntied = tr.run(5000)
    np.testing.assert_array_equal(pairs, f["pairs"].astype(np.uint32))
    np.testing.assert_array_equal(counts, f["counts"])
    np.testing.assert_array_equal(ntied, f["ntied"])
    ids = tr.ids()
    assert len(ids) == int(f["n_ids"][0])
    lens = tr.lengths(5000)
    assert int(lens[0]) == sym.numel() and int(lens[-1]) == len(ids) and np.all(np.diff(lens.astype(np.int64)) < 0)
    crc = np.bitwise_xor.reduce(ids.astype(np.uint64) * (np.arange(len(ids), dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1)))
    assert int(crc) == int(f["ids_crc"][0])
    # the live histogram equals a recount of the final stream (get_stats, lib.rs:28-48)
    h = tr.histogram()
    keys = (ids[:-1].astype(np.uint64) << np.uint64(32)) | ids[1:].astype(np.uint64)
    uk, uc = np.unique(keys, return_counts=True)
    want = {(int(k >> np.uint64(32)), int(k & np.uint64(0xFFFFFFFF))): int(c) for k, c in zip(uk, uc)}
    assert h == want


def _check_sharded(oracle, text, cuts, m, use_graph=True):
    from ecgbyte.dist_train import train_shards_local
    text = np.ascontiguousarray(text, np.uint8)
    bounds = [0] + list(cuts) + [len(text)]
    shards = [text[bounds[r]:bounds[r + 1]].tobytes() for r in range(len(bounds) - 1)]
    res, trs = train_shards_local(shards, m, use_graph=use_graph)  # graph: one captured step, replayed
    o_ids, o_pairs, o_counts, o_ntied = oracle.train_pairs(text, m, fast=len(text) > 20000)
    for pairs, counts, ntied in res:  # every rank reports the same merges
        np.testing.assert_array_equal(pairs, o_pairs)
        np.testing.assert_array_equal(counts, o_counts)
        np.testing.assert_array_equal(ntied, o_ntied)
    ids = np.concatenate([t.ids() for t in trs])
    np.testing.assert_array_equal(ids, o_ids)
    assert sum(int(t.lengths(len(o_pairs))[-1]) for t in trs) == len(o_ids)


def test_sharded_training_one_gpu_many_ranks(oracle, small_corpus):
    """The sharded kernels (halos, run parity across shards, delta lists, replicated
    histogram) driven by several trainers on one GPU == single-string training."""
    rng = np.random.default_rng(4)
    text = rng.integers(97, 100, size=5000).astype(np.uint8)
    _check_sharded(oracle, text, [1700, 3300], 60)
    _check_sharded(oracle, text, [1700, 3300], 60, use_graph=False)   # eager loop, same device-side step counter
    _check_sharded(oracle, text, [0, 1, 2, 4999], 40)          # empty and single-token shards
    runs = np.concatenate([np.full(4100, 105, np.uint8), rng.integers(104, 107, size=50).astype(np.uint8),
                           np.full(8300, 105, np.uint8), np.array([106, 105, 105], np.uint8)])
    for cuts in ([4096], [4097, 4150], [100, 4200, 12000], [6000, 6001, 6002]):
        _check_sharded(oracle, runs, cuts, 16)                  # (x,x) runs across shard and tile edges
    x, pct = small_corpus
    sym = oracle.quantize(x[:4], pct["percentile_1"], pct["percentile_99"]).reshape(-1)
    n = sym.size
    _check_sharded(oracle, sym, [n // 4, n // 2, 3 * n // 4], 300)


def _check_persistent(oracle, text, cuts, m, max_ctas=None):
    """dist_loop_kernel: one persistent cooperative kernel per rank, all co-resident on this GPU, exchanging
    patches / shard records through each other's receive areas (the NVLink protocol with plain pointers)."""
    from ecgbyte.dist_train import train_shards_persistent_local
    text = np.ascontiguousarray(text, np.uint8)
    bounds = [0] + list(cuts) + [len(text)]
    shards = [text[bounds[r]:bounds[r + 1]].tobytes() for r in range(len(bounds) - 1)]
    res, trs = train_shards_persistent_local(shards, m, max_ctas=max_ctas or 0)
    o_ids, o_pairs, o_counts, o_ntied = oracle.train_pairs(text, m, fast=len(text) > 20000)
    for pairs, counts, ntied in res:  # every rank reports the same merges
        np.testing.assert_array_equal(pairs, o_pairs)
        np.testing.assert_array_equal(counts, o_counts)
        np.testing.assert_array_equal(ntied, o_ntied)
    ids = np.concatenate([t.ids() for t in trs])
    np.testing.assert_array_equal(ids, o_ids)
    assert sum(int(t.lengths(len(o_pairs))[-1]) for t in trs) == len(o_ids)


def test_persistent_sharded_loop_one_gpu_many_ranks(oracle, small_corpus):
    rng = np.random.default_rng(4)
    text = rng.integers(97, 100, size=5000).astype(np.uint8)
    _check_persistent(oracle, text, [1700, 3300], 60)
    _check_persistent(oracle, text, [0, 1, 2, 4999], 40, max_ctas=8)   # empty and single-token shards
    _check_persistent(oracle, text, [1000, 2000, 3000, 4000], 30, max_ctas=2)   # fewer CTAs than peers: a CTA applies several lists
    runs = np.concatenate([np.full(4100, 105, np.uint8), rng.integers(104, 107, size=50).astype(np.uint8),
                           np.full(8300, 105, np.uint8), np.array([106, 105, 105], np.uint8)])
    for cuts in ([4096], [4097, 4150], [100, 4200, 12000], [6000, 6001, 6002]):
        _check_persistent(oracle, runs, cuts, 16, max_ctas=4)    # (x,x) runs across shards, chunks and tile edges
    x, pct = small_corpus
    sym = oracle.quantize(x[:4], pct["percentile_1"], pct["percentile_99"]).reshape(-1)
    n = sym.size
    _check_persistent(oracle, sym, [n // 4, n // 2, 3 * n // 4], 300)
    _check_persistent(oracle, sym, [n // 3], 300, max_ctas=2)    # few CTAs: several tiles per chunk


def test_persistent_sharded_loop_streaming_then_resident(oracle):
    """A shard too large for the CTAs' shared memory streams through HBM first (look-back pass, halo from the
    peers' records) and switches to the resident tail later, each rank at its own step."""
    from ecgbyte import synth
    x = synth.corpus(7, 6, L=5000, dtype=np.float32)
    pct = synth.percentiles(x, seed=1)
    sym = oracle.quantize(x, pct["percentile_1"], pct["percentile_99"]).reshape(-1)
    n = sym.size  # 360 000 symbols; 2 CTAs x 12 288 resident tokens per rank
    _check_persistent(oracle, sym, [n // 2 - 7], 120, max_ctas=2)
    _check_persistent(oracle, sym, [n // 5, n // 2], 120, max_ctas=3)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (CUDA IPC between one-process-per-GPU ranks)")
@pytest.mark.parametrize("mode", ["persistent", "auto"])
def test_sharded_training_real_ranks(mode):
    """tests/dist_gpu_train.py under torchrun on two GPUs: dist_loop_kernel with the peers' areas mapped through CUDA
    IPC (persistent) and the public auto policy; merge list, counts, tie log and merged stream == oracle."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(here, "dist_gpu_train.py"), "40", "300", mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "PARITY OK" in out.stdout
