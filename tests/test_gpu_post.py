"""GPU parity of the rows either side of the encoder (decode, dequantise, sequence packing)
against golden vectors from the reference's own code and against the Python restatement."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "post_reference.npz")


def test_decode_and_dequantize_match_reference(oracle):
    from ecgbyte.api import Vocab, dequantize
    g = np.load(GOLD)
    v = Vocab.from_pairs(g["dec_pairs"])
    pct = {"percentile_1": g["dec_pct"][0], "percentile_99": g["dec_pct"][1]}
    toks = [g["dec_tokens_%d" % r] for r in range(3)]
    stride = max(len(t) for t in toks)
    tok = np.zeros((3, stride), np.int32)
    for r, t in enumerate(toks):
        tok[r, : len(t)] = t
    lens = np.array([len(t) for t in toks], np.int32)
    sym, sym_len = v.decode_symbols(torch.from_numpy(tok).cuda(), torch.from_numpy(lens).cuda(), 12 * 400)
    for r in range(3):
        assert int(sym_len[r]) == len(g["dec_text_%d" % r])
        np.testing.assert_array_equal(sym[r, : int(sym_len[r])].cpu().numpy(), g["dec_text_%d" % r])
        vals = dequantize(sym[r], pct).cpu().numpy().reshape(g["dec_values_%d" % r].shape)
        np.testing.assert_array_equal(vals, g["dec_values_%d" % r])   # float64, bit-exact
    # encode -> decode round trip on the device (train_tokenizer.py:58-60)
    text = torch.from_numpy(np.stack([g["dec_text_%d" % r] for r in range(3)])).cuda()
    t2, l2 = v.encode_symbols(text)
    s2, sl2 = v.decode_symbols(t2, l2, text.shape[1])
    assert torch.equal(s2, text) and sl2.tolist() == [text.shape[1]] * 3
    # unknown token id -> loud error
    bad = torch.full((1, 4), 60000, dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError):
        v.decode_symbols(bad, torch.tensor([4], dtype=torch.int32, device="cuda"), 16)


def _pack_one(sig_ids, q, a, cfg):
    """sig_ids are LLM ids already; use an identity-like LUT over small fake token ids."""
    from ecgbyte.api import pack_training
    pad_to_max, pad_id, bos_id, eos_id, s0, s1 = cfg
    lut = torch.tensor(sig_ids if len(sig_ids) else [0], dtype=torch.int64)
    tokens = torch.arange(max(len(sig_ids), 1), dtype=torch.int32, device="cuda").unsqueeze(0)
    lens = torch.tensor([len(sig_ids)], dtype=torch.int32, device="cuda")
    out = pack_training(tokens, lens, lut, torch.tensor(q + a, dtype=torch.int64), torch.tensor([0, len(q) + len(a)]),
                        torch.tensor([len(q)], dtype=torch.int32), pad_to_max, pad_id, bos_id, eos_id, s0, s1)
    return [o[0].cpu().numpy() for o in out[:4]], int(out[4][0])


def test_pack_training_matches_reference():
    g = np.load(GOLD)
    for k in range(int(g["pack_n"][0])):
        cfg = g["pack_cfg_%d" % k].tolist()
        (ids, attn, labels, pos), status = _pack_one(g["pack_sig_%d" % k].tolist(), g["pack_q_%d" % k].tolist(),
                                                     g["pack_a_%d" % k].tolist(), cfg)
        assert status == 0
        np.testing.assert_array_equal(ids, g["pack_ids_%d" % k])
        np.testing.assert_array_equal(attn, g["pack_attn_%d" % k])
        np.testing.assert_array_equal(labels, g["pack_labels_%d" % k])
        np.testing.assert_array_equal(pos, g["pack_pos_%d" % k])


def test_pack_training_random_vs_restatement():
    from oracle import py_restatement as P
    rng = np.random.default_rng(9)
    for _ in range(20):
        pad_to_max = int(rng.integers(8, 1100))
        nq, na = int(rng.integers(0, pad_to_max // 2 + 1)), int(rng.integers(0, pad_to_max // 2 + 1))
        nsig = int(rng.integers(0, 2 * pad_to_max))
        sig = rng.integers(1000, 2000, size=nsig).tolist()
        q, a = rng.integers(0, 900, size=nq).tolist(), rng.integers(0, 900, size=na).tolist()
        cfg = [pad_to_max, 999999, 5, 6, 7, 8]
        (ids, attn, labels, pos), status = _pack_one(sig, q, a, cfg)
        w = P.prepare_training(sig, q, a, *cfg)
        assert status == 0
        for got, want in zip((ids, attn, labels, pos), w):
            np.testing.assert_array_equal(got, want)
    # question + answer longer than pad_to_max: flagged (the reference's assert fails there)
    _, status = _pack_one([1, 2, 3], list(range(30)), list(range(30)), [40, 999999, 5, 6, 7, 8])
    assert status == 1


def test_batcher_end_to_end(oracle, small_corpus, small_table):
    """records -> fused encode -> pack == reference pipeline restated on the CPU."""
    from ecgbyte.data_loader import ECGTokenBatcher
    from oracle import py_restatement as P
    x, pct = small_corpus
    _, vocab, merges = small_table
    lut = torch.arange(len(vocab), dtype=torch.int64) + 128257   # 'signal_k' -> 128257 + k
    b = ECGTokenBatcher(merges, pct, lut, pad_to_max=1020, pad_id=128256, bos_id=128000, eos_id=128001,
                        sig_start_id=140000, sig_end_id=140001, dtype=torch.float64)
    qs = [[11, 12, 13], [21, 22], [31], [41, 42, 43, 44]]
    ans = [[5, 6], [7], [8, 9, 10], [1]]
    out = b(x[:4], qs, ans)
    trie = oracle.Trie(merges=merges)
    for r in range(4):
        sym = oracle.quantize(x[r], pct["percentile_1"], pct["percentile_99"]).reshape(-1)
        sig = (trie.encode(sym).astype(np.int64) + 128257).tolist()
        w = P.prepare_training(sig, qs[r], ans[r], 1020, 128256, 128000, 128001, 140000, 140001)
        np.testing.assert_array_equal(out["tokenized_signal"][r].cpu().numpy(), w[0])
        np.testing.assert_array_equal(out["attn_mask"][r].cpu().numpy(), w[1])
        np.testing.assert_array_equal(out["quantized_signal_ids_input"][r].cpu().numpy(), w[2])
        np.testing.assert_array_equal(out["position_ids"][r].cpu().numpy(), w[3])


def test_global_stats_match_numpy(small_corpus):
    """compute_global_stats' arithmetic (preprocess_utils.py:183-206): np.min / np.max over the segments and
    np.percentile(samples, 1 / 99) -- NumPy is the reference's own implementation of these."""
    from ecgbyte.api import global_stats, minmax, percentiles
    rng = np.random.default_rng(21)
    x, _ = small_corpus
    for dt in (np.float64, np.float32):
        seg = x.astype(dt)
        lo, hi = minmax(torch.from_numpy(seg).cuda())
        assert lo == float(seg.min()) and hi == float(seg.max())
    xi = rng.integers(-30000, 30000, size=100001).astype(np.int16)
    assert minmax(torch.from_numpy(xi).cuda()) == (float(xi.min()), float(xi.max()))
    for n in (1, 2, 3, 100, 100000, 100003):
        s = rng.normal(0.2, 0.7, size=n)
        if n > 50:
            s[::17] = s[5]          # ties
        qs = [0, 1, 25, 50, 99, 100, 33.3]
        got = percentiles(torch.from_numpy(s).cuda(), qs)
        want = np.percentile(s, qs)
        np.testing.assert_array_equal(got, want)      # bit-exact float64
    # the reference's sample: all values of the first segments, then a random subset (pu.py:190-195)
    flat = x.reshape(-1)
    sample = np.concatenate([flat[:90000], rng.choice(flat[90000:120000], 10000, replace=False)])
    st = global_stats(torch.from_numpy(x).cuda(), torch.from_numpy(sample).cuda())
    assert st["percentile_1"] == np.percentile(sample, 1) and st["percentile_99"] == np.percentile(sample, 99)
    assert st["global_min"] == x.min() and st["global_max"] == x.max()
    # NaN propagates like np.min / np.percentile
    bad = s.copy(); bad[7] = np.nan
    assert np.isnan(percentiles(torch.from_numpy(bad).cuda(), [1])[0])
    assert all(np.isnan(v) for v in minmax(torch.from_numpy(bad).cuda()))


def _vocab_strings(pairs):
    vocab = {i: (chr(i) if i <= 127 else "<%d>" % i) for i in range(256)}
    for k, (l, r) in enumerate(pairs):
        vocab[256 + k] = vocab[int(l)] + vocab[int(r)]
    return vocab


def test_expand_attention_matches_restatement():
    """runners/interpret.py:106-111 on the device vs the line-by-line restatement."""
    from oracle import py_restatement as P
    from ecgbyte.api import Vocab
    from ecgbyte import tokenizer_utils as tu
    g = np.load(GOLD)
    v = Vocab.from_pairs(g["dec_pairs"])
    vocab = _vocab_strings(g["dec_pairs"])
    toks = [g["dec_tokens_%d" % r] for r in range(3)]
    stride = max(len(t) for t in toks)
    rng = np.random.default_rng(5)
    tok = np.zeros((3, stride), np.int32)
    att = rng.random((3, stride)).astype(np.float32)
    for r, t in enumerate(toks):
        tok[r, : len(t)] = t
    lens = np.array([len(t) for t in toks], np.int32)
    out, out_len = v.expand_attention(torch.from_numpy(tok).cuda(), torch.from_numpy(lens).cuda(),
                                      torch.from_numpy(att).cuda(), 12 * 400)
    for r in range(3):
        want = np.array(P.expand_attention(toks[r].tolist(), att[r, : lens[r]].tolist(), vocab), np.float32)
        assert int(out_len[r]) == len(want) == len(g["dec_text_%d" % r])
        np.testing.assert_array_equal(out[r, : len(want)].cpu().numpy(), want)
    # the drop-in signature, host path and device path
    ids, a = toks[0].tolist(), att[0, : lens[0]].tolist()
    assert tu.expand_attention(ids, a, vocab) == P.expand_attention(ids, a, vocab)
    merges = [(list(vocab[256 + k].encode()), 256 + k) for k in range(len(g["dec_pairs"]))]
    np.testing.assert_array_equal(np.array(tu.expand_attention(ids, a, vocab, merges), np.float32),
                                  np.array(P.expand_attention(ids, a, vocab), np.float32))
    # empty record, unknown id
    z, zl = v.expand_attention(torch.zeros((1, 4), dtype=torch.int32, device="cuda"),
                               torch.zeros((1,), dtype=torch.int32, device="cuda"),
                               torch.zeros((1, 4), device="cuda"), 8)
    assert int(zl[0]) == 0
    with pytest.raises(ValueError):
        v.expand_attention(torch.full((1, 2), 60000, dtype=torch.int32, device="cuda"),
                           torch.tensor([2], dtype=torch.int32, device="cuda"), torch.ones((1, 2), device="cuda"), 8)


def test_token_histogram_matches_counter(tmp_path):
    """analyze_token_distribution (tokenizer_utils.py:30-54): Counter of encoded ids + lengths."""
    from oracle import py_restatement as P
    from ecgbyte import synth
    from ecgbyte.api import Quantizer, Vocab, token_histogram
    from ecgbyte import tokenizer_utils as tu
    g = np.load(GOLD)
    pairs = g["dec_pairs"]
    vocab = _vocab_strings(pairs)
    merges = [(list(vocab[256 + k].encode()), 256 + k) for k in range(len(pairs))]
    pct = {"percentile_1": float(g["dec_pct"][0]), "percentile_99": float(g["dec_pct"][1])}
    recs = [synth.record(3, k, 400).astype(np.float64) for k in range(5)]
    paths = []
    for k, r in enumerate(recs):
        paths.append(str(tmp_path / ("r%d.npy" % k)))
        np.save(paths[-1], r)
    want_ids = [P.encode_text(bytes(P.normalize_all_symbols(r, pct["percentile_1"], pct["percentile_99"]).reshape(-1)), merges)
                for r in recs]
    want_counts, want_lengths = P.token_distribution(want_ids)
    counts, lengths = tu.analyze_token_distribution(paths, merges, pct, num_workers=2, batch=2)
    assert lengths == want_lengths
    assert counts == want_counts
    # the primitive: shared-memory counters and (large id space) global atomics give the same histogram
    v = Vocab(merges)
    q = Quantizer(pct, dtype=torch.float64)
    tokens, lens = v.encode_batch(q, torch.from_numpy(np.stack(recs)).cuda())
    n_ids = 256 + len(pairs)
    h1 = token_histogram(tokens, lens, n_ids).cpu().numpy()
    h2 = token_histogram(tokens, lens, 40000).cpu().numpy()
    assert {int(i): int(h1[i]) for i in np.nonzero(h1)[0]} == dict(want_counts)
    np.testing.assert_array_equal(h2[:n_ids], h1)
    assert h2[n_ids:].sum() == 0
    with pytest.raises(ValueError):
        token_histogram(tokens, lens, 100)   # ids >= 100 occur


def test_encode_pipeline_csr(oracle, small_corpus, small_table):
    """Host-buffer pipeline with compact output (2-byte ids stored by the kernel straight into pinned host memory, rows
    back to back inside a chunk) == oracle, over several chunks and two pipelined calls, and == the padded int32 pipeline."""
    from ecgbyte.api import EncodePipeline, EncodePipelineCSR, Quantizer, Vocab
    x, pct = small_corpus
    _, vocab, merges = small_table
    n = x.shape[0]
    q = Quantizer(pct, dtype=torch.float32, device="cuda:0")
    v = Vocab(merges, device="cuda:0")
    rec_len = x.shape[1] * x.shape[2]
    stride = 2048
    xh = torch.from_numpy(x.astype(np.float32)).pin_memory()
    tok16 = torch.zeros(n * stride, dtype=torch.uint16).pin_memory()
    lens = torch.zeros(n, dtype=torch.int32).pin_memory()
    offp = torch.zeros(n, dtype=torch.int64).pin_memory()
    pipe = EncodePipelineCSR(v, q, rec_len, stride, chunk=3, depth=2)   # ragged last chunk, slots reused
    for _ in range(2):                                                  # consecutive calls pipeline into each other
        pipe.run(xh, tok16, lens, offp)
    torch.cuda.synchronize()
    off = offp.numpy()
    trie = oracle.Trie(merges=merges)
    t16 = tok16.numpy()
    want_off = 0
    for r in range(n):
        if r % 3 == 0:
            want_off = r * stride                                        # chunk starts at a fixed place
        assert off[r] == want_off                                        # rows back to back inside the chunk
        sym = oracle.quantize(x[r].astype(np.float32), pct["percentile_1"], pct["percentile_99"]).reshape(-1)
        want = trie.encode(sym)
        np.testing.assert_array_equal(t16[off[r]:off[r] + lens[r]].astype(np.uint32), want.astype(np.uint32))
        want_off += len(want)
    tok32 = torch.zeros((n, stride), dtype=torch.int32).pin_memory()
    lens2 = torch.zeros(n, dtype=torch.int32).pin_memory()
    EncodePipeline(v, q, rec_len, stride, chunk=3, depth=2).run(xh, tok32, lens2)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(lens.numpy(), lens2.numpy())
    for r in range(n):
        np.testing.assert_array_equal(t16[off[r]:off[r] + lens[r]].astype(np.int32), tok32[r, : lens2[r]].numpy())
