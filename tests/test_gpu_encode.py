"""K2 parity: CUDA greedy longest-match encoder vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _check_batch(oracle, vocab, trie, sym):
    want_tok, want_len = trie.encode_batch(sym)
    tok, lens = vocab.encode_symbols(torch.from_numpy(sym).cuda())
    tok, lens = tok.cpu().numpy(), lens.cpu().numpy()
    np.testing.assert_array_equal(lens, want_len.astype(np.int32))
    for r in range(sym.shape[0]):
        np.testing.assert_array_equal(tok[r, : lens[r]], want_tok[r, : want_len[r]].astype(np.int32))


def test_encode_symbols_matches_oracle(oracle, small_corpus, small_table):
    from ecgbyte.api import Vocab
    x, pct = small_corpus
    _, _, merges = small_table
    sym = oracle.quantize(x, pct["percentile_1"], pct["percentile_99"]).reshape(x.shape[0], -1)
    v = Vocab(merges=merges)
    info = v.info()
    assert info["compact"] == 1 and info["n_merges"] == len(merges)
    trie = oracle.Trie(merges=merges)
    assert info["n_nodes"] == trie.nodes - (256 - 26)  # bytes outside a..z stay implicit
    _check_batch(oracle, v, trie, sym)


@pytest.mark.parametrize("dtype", ["float32", "float64", "int16"])
def test_fused_encode_matches_oracle(oracle, small_corpus, small_table, dtype):
    from ecgbyte import synth
    from ecgbyte.api import Quantizer, Vocab
    x64, pct = small_corpus
    _, _, merges = small_table
    x = synth.cast(x64, np.dtype(dtype))
    sym = oracle.quantize(x, pct["percentile_1"], pct["percentile_99"]).reshape(x.shape[0], -1)
    trie = oracle.Trie(merges=merges)
    want_tok, want_len = trie.encode_batch(sym)
    v = Vocab(merges=merges)
    q = Quantizer(pct, dtype=getattr(torch, dtype))
    tok, lens = v.encode_batch(q, torch.from_numpy(x).cuda())
    tok, lens = tok.cpu().numpy(), lens.cpu().numpy()
    np.testing.assert_array_equal(lens, want_len.astype(np.int32))
    for r in range(x.shape[0]):
        np.testing.assert_array_equal(tok[r, : lens[r]], want_tok[r, : want_len[r]].astype(np.int32))
    # host-buffer entry point, truncated output: lengths stay true, stored prefix matches
    tok_h, len_h = v.encode_batch_host(q, x, 64)
    np.testing.assert_array_equal(len_h, want_len.astype(np.int32))
    for r in range(x.shape[0]):
        m = min(64, want_len[r])
        np.testing.assert_array_equal(tok_h[r, :m], want_tok[r, :m].astype(np.int32))


def test_encode_kats():
    """Hand-derived known answers."""
    import rust_bpe
    # longest match, not merge order: 'abc' with [bc->256, ab->257] -> [257, 'c']
    assert rust_bpe.encode_text("abc", [([98, 99], 256), ([97, 98], 257)]) == [257, 99]
    # a duplicate sequence: the later id wins (lib.rs:145)
    assert rust_bpe.encode_text("abab", [([97, 98], 256), ([97, 98], 300)]) == [300, 300]
    # non-terminal interior node: 'aaaa' is a token, 'aaa' is not
    m = [([97, 97], 256), ([97, 97, 97, 97], 257)]
    assert rust_bpe.encode_text("aaa", m) == [256, 97]
    assert rust_bpe.encode_text("aaaaa", m) == [257, 97]
    assert rust_bpe.encode_text("", m) == []
    # bytes that occur in no merge, and non-ASCII text (UTF-8 bytes, lib.rs:151)
    assert rust_bpe.encode_text("a-b", m) == [97, 45, 98]
    assert rust_bpe.encode_text("é", []) == [0xC3, 0xA9]
    with pytest.raises(TypeError):
        rust_bpe.encode_text(b"abc", m)
    with pytest.raises(TypeError):
        rust_bpe.encode_text("abc", "nope")


def test_encode_ragged_and_edge_records(oracle, small_table):
    from ecgbyte.api import Vocab
    _, _, merges = small_table
    v = Vocab(merges=merges)
    trie = oracle.Trie(merges=merges)
    rng = np.random.default_rng(5)
    lens = [0, 1, 2, 7, 8, 9, 15, 16, 17, 33, 1000, 0, 4099]
    parts = [rng.integers(97, 123, size=n).astype(np.uint8) for n in lens]
    # long runs and a constant record exercise deep walks and long reach-backs
    parts[10][:] = 105
    parts[12][:2000] = 104
    flat = np.concatenate(parts + [np.zeros(1, np.uint8)])[:-1]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    tok, ln = v.encode_symbols(torch.from_numpy(flat).cuda(), offsets=torch.from_numpy(off))
    tok, ln = tok.cpu().numpy(), ln.cpu().numpy()
    for r, p in enumerate(parts):
        want = trie.encode(p)
        assert ln[r] == len(want)
        np.testing.assert_array_equal(tok[r, : ln[r]], want.astype(np.int32))


def test_encode_wide_alphabet(oracle):
    """> 31 symbol classes -> wide-node kernel (general text)."""
    import rust_bpe
    text = "".join(chr(33 + (i * 7 + (i // 5) * 3) % 60) for i in range(4000)) * 2
    ids, vocab, merges = oracle.byte_pair_encoding(text, 200, fast=True)
    from ecgbyte.api import Vocab
    assert Vocab(merges=merges).info()["compact"] == 0
    assert rust_bpe.encode_text(text, merges) == oracle.encode_text(text, merges)


def test_roundtrip_decode(oracle, small_corpus, small_table):
    from ecgbyte import tokenizer_utils as tu
    x, pct = small_corpus
    _, vocab, merges = small_table
    s = tu.process_ecg(x[3], pct)
    ids = tu.encode_text(s, merges)
    assert tu.decode_text(ids, vocab) == s  # train_tokenizer.py:58-60
    assert ids == oracle.encode_text(s, merges)


def test_encode_large_vocab_partial_smem(oracle):
    """A vocabulary whose trie (53k nodes = 430 KB) does not fit in shared memory: the leading
    nodes are staged, the rest is read through L1/L2 (ALL_SMEM = false path)."""
    import itertools
    from ecgbyte.api import Quantizer, Vocab
    letters = list(range(97, 123))
    seqs = [list(p) for p in itertools.product(letters, repeat=2)]
    seqs += [list(p) for p in itertools.product(letters, repeat=3)]
    seqs += [[f] + list(p) for f in (97, 98) for p in itertools.product(letters, repeat=3)]
    merges = [(s, 256 + i) for i, s in enumerate(seqs)]
    v = Vocab(merges=merges)
    info = v.info()
    assert info["compact"] == 1 and info["n_nodes"] > 50000
    trie = oracle.Trie(merges=merges)
    rng = np.random.default_rng(12)
    sym = rng.integers(97, 123, size=(40, 3000)).astype(np.uint8)
    sym[:, ::7] = 97  # plenty of a/b-led 4-grams
    _check_batch(oracle, v, trie, sym)
    # fused path with the same vocabulary
    x = rng.normal(0.3, 0.4, size=(8, 12, 250)).astype(np.float32)
    pct = {"percentile_1": -0.2, "percentile_99": 0.9}
    q = Quantizer(pct)
    s2 = oracle.quantize(x, -0.2, 0.9).reshape(8, -1)
    w_tok, w_len = trie.encode_batch(s2)
    tok, lens = v.encode_batch(q, torch.from_numpy(x).cuda())
    np.testing.assert_array_equal(lens.cpu().numpy(), w_len.astype(np.int32))
    for r in range(8):
        np.testing.assert_array_equal(tok[r, : w_len[r]].cpu().numpy(), w_tok[r, : w_len[r]].astype(np.int32))


def test_encode_many_records_multiple_passes(oracle, small_table):
    """More records than walkers (CTAs loop over their range) and out_stride truncation."""
    from ecgbyte.api import Vocab
    _, _, merges = small_table
    v = Vocab(merges=merges)
    trie = oracle.Trie(merges=merges)
    rng = np.random.default_rng(13)
    n = 148 * 768 + 1000
    sym = rng.integers(104, 108, size=(n, 48)).astype(np.uint8)
    tok, lens = v.encode_symbols(torch.from_numpy(sym).cuda(), out_stride=8)
    tok, lens = tok.cpu().numpy(), lens.cpu().numpy()
    idx = rng.choice(n, size=300, replace=False)
    w_tok, w_len = trie.encode_batch(sym[idx])
    np.testing.assert_array_equal(lens[idx], w_len.astype(np.int32))   # true counts even when truncated
    for k, r in enumerate(idx):
        m = min(8, w_len[k])
        np.testing.assert_array_equal(tok[r, :m], w_tok[k, :m].astype(np.int32))


def test_full_size_roundtrip_100k_records():
    """BASELINE.json config 2 at full size (100k records x 12 x 5000 fp32 = 24 GB, 5,000-merge
    table), checked through size-independent properties on the device:
      decode(encode(x)) == quantise(x) for every record, token counts consistent, and the fused
      encoder == quantise-then-encode-symbols; plus a 32-record sample against the CPU oracle."""
    import os
    from ecgbyte import synth
    from ecgbyte.api import Quantizer, Vocab
    from oracle import oracle as O
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs ~45 GB of device memory")
    f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ptbxl_1000_m5000.npz"))
    pct = {"percentile_1": f["pct"][0], "percentile_99": f["pct"][1]}
    pairs = f["pairs"].astype(np.uint32)
    v = Vocab.from_pairs(pairs)
    q = Quantizer(pct, dtype=torch.float32)
    n, stride = 100000, 8192
    x = synth.corpus_cuda(77, n, 5000, torch.float32, "cuda:0")
    tok, lens = v.encode_batch(q, x, out_stride=stride)
    assert int(lens.max()) <= stride and int(lens.min()) > 0
    total_sym = 0
    for c0 in range(0, n, 20000):   # chunked to bound the symbol buffers
        sl = slice(c0, c0 + 20000)
        sym = q.quantize(x[sl]).reshape(20000, -1)
        dec, dec_len = v.decode_symbols(tok[sl], lens[sl], sym.shape[1])
        assert bool((dec_len == sym.shape[1]).all())
        assert torch.equal(dec, sym)
        t2, l2 = v.encode_symbols(sym, out_stride=stride)
        assert torch.equal(l2, lens[sl])
        m = torch.arange(stride, device="cuda").unsqueeze(0) < l2.unsqueeze(1)
        assert torch.equal(torch.where(m, t2, 0), torch.where(m, tok[sl], 0))
        total_sym += int(dec_len.sum())
    assert total_sym == n * 60000
    # every token id is a valid vocabulary id
    m_all = torch.arange(stride, device="cuda").unsqueeze(0) < lens.unsqueeze(1)
    assert int(torch.where(m_all, tok, 0).max()) < 256 + len(pairs)
    idx = np.random.default_rng(1).choice(n, size=32, replace=False)
    seq, off = O.expand(pairs)
    trie = O.Trie(flat=(seq, off, np.arange(256, 256 + len(pairs), dtype=np.uint32)))
    xs = x[torch.from_numpy(idx).cuda()].cpu().numpy()
    w_tok, w_len = trie.encode_batch(O.quantize(xs, pct["percentile_1"], pct["percentile_99"]).reshape(32, -1), stride)
    g_tok, g_len = tok[torch.from_numpy(idx).cuda()].cpu().numpy(), lens[torch.from_numpy(idx).cuda()].cpu().numpy()
    np.testing.assert_array_equal(g_len, w_len.astype(np.int32))
    for k in range(32):
        np.testing.assert_array_equal(g_tok[k, : w_len[k]], w_tok[k, : w_len[k]].astype(np.int32))


def test_long_string_position_parallel(oracle, small_corpus, small_table):
    """rust_bpe.encode_text on long strings takes the position-parallel path (encode_long.cu);
    it must equal the sequential greedy longest match."""
    import rust_bpe
    x, pct = small_corpus
    _, vocab, merges = small_table
    sym = oracle.quantize(x, pct["percentile_1"], pct["percentile_99"]).reshape(-1)
    rng = np.random.default_rng(3)
    cases = [sym[:4096], sym[:4097], sym[:60000], sym,                       # ECG text, chunk edges, whole corpus
             np.full(70001, 105, np.uint8),                                   # one run: tokens longer than walks
             rng.integers(97, 123, size=50000).astype(np.uint8)]              # incompressible
    mixed = sym[:30000].copy()
    mixed[::97] = 45                                                          # bytes that occur in no merge
    cases.append(mixed)
    trie = oracle.Trie(merges=merges)
    for s in cases:
        got = rust_bpe.encode_text(s.tobytes().decode("latin1") if s.max() < 128 else None, merges)
        want = trie.encode(s)
        assert len(got) == len(want)
        np.testing.assert_array_equal(np.array(got, np.uint32), want)


def test_encode_10k_merge_table_pair_walker(oracle):
    """BASELINE config 3's table (10,000 merges, oracle fixture): its 8-byte trie nodes do not fit in shared memory next
    to the walkers' rings, so the fused encoder dispatches to the two-symbol-stride pair-table walker (encode2.cu).
    fp32, int16 and text input, ragged record count; == oracle."""
    import os
    from ecgbyte import synth
    from ecgbyte.api import Quantizer, Vocab
    f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ptbxl_1000_m10000.npz"))
    pairs = f["pairs"].astype(np.uint32)
    pct = {"percentile_1": np.float64(f["pct"][0]), "percentile_99": np.float64(f["pct"][1])}
    v = Vocab.from_pairs(pairs, device="cuda:0")
    info = v.info()
    assert info["pair_slots"] > 0 and info["n_nodes"] * 8 + 768 * 128 > 227 * 1024
    x = synth.corpus(99, 37, L=5000, dtype=np.float32)
    sym = oracle.quantize(x, pct["percentile_1"], pct["percentile_99"]).reshape(37, -1)
    seq, off = oracle.expand(pairs)
    trie = oracle.Trie(flat=(seq, off, np.arange(256, 256 + len(pairs), dtype=np.uint32)))
    w_tok, w_len = trie.encode_batch(sym, 8192)
    q = Quantizer(pct, dtype=torch.float32, device="cuda:0")
    tok, lens = v.encode_batch(q, torch.from_numpy(x).cuda(), out_stride=8192)
    tok, lens = tok.cpu().numpy(), lens.cpu().numpy()
    np.testing.assert_array_equal(lens, w_len.astype(np.int32))
    for r in range(37):
        np.testing.assert_array_equal(tok[r, : lens[r]], w_tok[r, : w_len[r]].astype(np.int32))
    # the same symbols as text
    tok2, lens2 = v.encode_symbols(torch.from_numpy(sym).cuda(), out_stride=8192)
    np.testing.assert_array_equal(lens2.cpu().numpy(), lens)
    np.testing.assert_array_equal(tok2.cpu().numpy()[3, : lens[3]], tok[3, : lens[3]])
    # int16 records (explicit 1e-3 de-scaling API)
    x16 = np.clip(np.round(x * 1000.0), -32768, 32767).astype(np.int16)
    sym16 = oracle.quantize(x16, pct["percentile_1"], pct["percentile_99"]).reshape(37, -1)
    w16, l16 = trie.encode_batch(sym16, 8192)
    q16 = Quantizer(pct, dtype=torch.int16, device="cuda:0")
    t16, n16 = v.encode_batch(q16, torch.from_numpy(x16).cuda(), out_stride=8192)
    np.testing.assert_array_equal(n16.cpu().numpy(), l16.astype(np.int32))
    np.testing.assert_array_equal(t16.cpu().numpy()[36, : l16[36]], w16[36, : l16[36]].astype(np.int32))
