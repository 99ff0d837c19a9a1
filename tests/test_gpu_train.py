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
