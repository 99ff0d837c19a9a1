"""CPU: the Python restatement of the rows either side of the encoder against golden vectors
produced by the reference's own code (oracle/make_golden_post.py)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "post_reference.npz")


def test_prepare_training_restatement_matches_reference():
    from oracle import py_restatement as P
    g = np.load(GOLD)
    for k in range(int(g["pack_n"][0])):
        pad_to_max, pad_id, bos_id, eos_id, s0, s1 = g["pack_cfg_%d" % k].tolist()
        ids, attn, labels, pos = P.prepare_training(g["pack_sig_%d" % k].tolist(), g["pack_q_%d" % k].tolist(),
                                                    g["pack_a_%d" % k].tolist(), pad_to_max, pad_id, bos_id, eos_id, s0, s1)
        np.testing.assert_array_equal(ids, g["pack_ids_%d" % k])
        np.testing.assert_array_equal(attn, g["pack_attn_%d" % k])
        np.testing.assert_array_equal(labels, g["pack_labels_%d" % k])
        np.testing.assert_array_equal(pos, g["pack_pos_%d" % k])


def test_decode_restatement_matches_reference(oracle):
    from oracle import py_restatement as P
    g = np.load(GOLD)
    _, _, merges = oracle.to_reference_types(np.zeros(0, np.uint32), g["dec_pairs"])
    p1, p99 = g["dec_pct"]
    for r in range(3):
        sym = P.decode_symbols(g["dec_tokens_%d" % r], merges)
        np.testing.assert_array_equal(sym, g["dec_text_%d" % r])
        np.testing.assert_array_equal(oracle.decode(g["dec_tokens_%d" % r], merges=merges), g["dec_text_%d" % r])
        vals = P.reverse_normalize_all(sym, p1, p99).reshape(g["dec_values_%d" % r].shape)
        np.testing.assert_array_equal(vals, g["dec_values_%d" % r])


def test_attention_and_distribution_restatements():
    """Known answers for runners/interpret.py:106-111 and tokenizer_utils.py:30-54 (Counter + lengths)."""
    from oracle import py_restatement as P
    vocab = {97: "a", 98: "b", 256: "ab", 257: "abab", 200: "<200>"}
    assert P.expand_attention([257, 97, 256], [0.5, 1.0, 2.0], vocab) == [0.5] * 4 + [1.0] + [2.0] * 2
    assert P.expand_attention([257, 97], [0.5], vocab) == [0.5] * 4          # zip stops at the shorter input
    assert P.expand_attention([200], [3.0], vocab) == [3.0] * 5              # characters of the vocab string
    counts, lengths = P.token_distribution([[257, 97, 257], [], [97]])
    assert counts == {257: 2, 97: 2} and lengths == [3, 0, 1]


def test_expand_attention_mirror_host_path():
    """The drop-in signature (runners/interpret.py:106-111) without merges runs on the host, like the reference."""
    from oracle import py_restatement as P
    from ecgbyte import tokenizer_utils as tu
    vocab = {97: "a", 98: "b", 256: "ab", 257: "abab"}
    ids, att = [257, 97, 256, 98], [0.25, 1.0, 2.0, 4.0]
    assert tu.expand_attention(ids, att, vocab) == P.expand_attention(ids, att, vocab)
    assert tu.expand_attention([], [], vocab) == []
