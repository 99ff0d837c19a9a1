"""Call sites of the hot path, CPU side: the oracle and the host logic of the mirrors against golden vectors made by
the reference's OWN Python (oracle/make_golden_callsites.py), and the (vocab, merges) pickle in both directions."""
import os
import pickle
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "callsites_reference.npz")
PKL = os.path.join(HERE, "golden", "ref_vocab_merges.pkl")
REF = "/root/reference"


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _pct(g):
    return {"percentile_1": np.float64(g["pct"][0]), "percentile_99": np.float64(g["pct"][1])}


def test_oracle_matches_reference_process_ecg(oracle, gold):
    """process_ecg (tu.py:56-59) per file incl. the float32 and the RAW-integer int16 record, and the joined corpus
    string of process_large_file (tu.py:79-93) in list order / with the n cap."""
    pct = _pct(gold)
    strings = []
    for i in range(6):
        r = gold["rec_%d" % i]
        x = r if r.dtype in (np.float32, np.float64) else r.astype(np.float64)   # NumPy promotes integer records
        s = oracle.quantize(x, pct["percentile_1"], pct["percentile_99"]).reshape(-1)
        np.testing.assert_array_equal(s, gold["pe_%d" % i])
        strings.append(s)
    order = gold["plf_order"]
    np.testing.assert_array_equal(np.concatenate([strings[i] for i in order]), gold["plf_all"])
    np.testing.assert_array_equal(np.concatenate([strings[i] for i in order[:4]]), gold["plf_n4"])


def test_oracle_matches_reference_token_distribution(oracle, gold):
    """analyze_token_distribution (tu.py:30-54) as the reference ran it around the encode call."""
    from collections import Counter
    pct = _pct(gold)
    _, vocab, merges = oracle.to_reference_types([], gold["pairs"])
    trie = oracle.Trie(merges=merges)
    counts, lengths = Counter(), []
    for i in gold["atd_files"]:
        r = gold["rec_%d" % i]
        x = r if r.dtype in (np.float32, np.float64) else r.astype(np.float64)
        ids = trie.encode(oracle.quantize(x, pct["percentile_1"], pct["percentile_99"]).reshape(-1))
        counts.update(int(t) for t in ids)
        lengths.append(len(ids))
    assert sorted(counts) == gold["atd_ids"].tolist()
    assert [counts[k] for k in sorted(counts)] == gold["atd_counts"].tolist()
    assert lengths == gold["atd_lengths"].tolist()


def test_expand_attention_host_path_matches_reference(gold):
    """runners/interpret.py:106-111 through the mirror's host path (no device needed), incl. raw bytes > 127 whose
    vocab string is 5 characters long, and zip() stopping at the shorter argument."""
    sys.modules.setdefault("torch", pytest.importorskip("torch"))
    from ecgbyte.tokenizer_utils import expand_attention
    with open(PKL, "rb") as f:
        vocab, merges = pickle.load(f)
    for k in (0, 1):
        got = expand_attention(gold["ea_ids_%d" % k].tolist(), gold["ea_att_%d" % k].tolist(), vocab)
        assert got == gold["ea_out_%d" % k].tolist()
    got = expand_attention(gold["ea_ids_0"].tolist()[:7], gold["ea_att_0"].tolist()[:4], vocab)
    assert got == gold["ea_out_short"].tolist()


def test_pickle_written_by_reference_loads_in_mirror(oracle, gold):
    """tests/golden/ref_vocab_merges.pkl was written by the reference's save_vocab_and_merges (tu.py:62-64)."""
    from ecgbyte.tokenizer_utils import load_vocab_and_merges
    vocab, merges = load_vocab_and_merges(PKL)
    _, o_vocab, o_merges = oracle.to_reference_types([], gold["pairs"])
    assert isinstance(vocab, dict) and isinstance(merges, list)
    assert vocab == o_vocab
    assert [(list(s), int(i)) for s, i in merges] == [(list(s), int(i)) for s, i in o_merges]
    assert all(isinstance(k, int) and isinstance(v, str) for k, v in vocab.items())
    assert all(isinstance(m, tuple) and isinstance(m[0], list) and isinstance(m[1], int) for m in merges)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree exists in the build container only")
def test_pickle_cross_load_with_reference_loader(tmp_path, oracle, gold):
    """Write with the mirror, load with the REFERENCE's load_vocab_and_merges (tu.py:66-69), and the reverse."""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
    from make_golden import import_reference_tu
    tu = import_reference_tu()
    from ecgbyte import tokenizer_utils as mirror
    _, vocab, merges = oracle.to_reference_types([], gold["pairs"])
    a = str(tmp_path / "mirror.pkl")
    mirror.save_vocab_and_merges(vocab, merges, a)
    v2, m2 = tu.load_vocab_and_merges(a)                 # the reference reads the mirror's file
    assert v2 == vocab and m2 == merges
    b = str(tmp_path / "reference.pkl")
    tu.save_vocab_and_merges(vocab, merges, b)            # the mirror reads the reference's file
    v3, m3 = mirror.load_vocab_and_merges(b)
    assert v3 == vocab and m3 == merges
    assert open(a, "rb").read() == open(b, "rb").read()   # byte-identical files
