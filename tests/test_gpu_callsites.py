"""The reference's call sites of the hot path through the drop-in (CUDA) against golden vectors produced by the
reference's OWN Python (oracle/make_golden_callsites.py): process_large_file / process_ecg (tu.py:56-59, 79-93),
analyze_token_distribution (tu.py:30-54), expand_attention (runners/interpret.py:106-111), and the call sequence of
train_tokenizer.py:19-45 + its self-check :47-64."""
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "callsites_reference.npz")
PKL = os.path.join(HERE, "golden", "ref_vocab_merges.pkl")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture()
def files(tmp_path, gold):
    paths = []
    for i in range(6):
        p = str(tmp_path / ("ecg_%d.npy" % i))
        np.save(p, gold["rec_%d" % i])
        paths.append(p)
    return paths


def _pct(g):
    return {"percentile_1": np.float64(g["pct"][0]), "percentile_99": np.float64(g["pct"][1])}


def test_process_large_file_matches_reference(tmp_path, gold, files):
    """File order, .strip() of padded lines, the n cap, mixed dtypes (float64, float32, RAW-integer int16) and a
    record of another shape in the middle of a batch."""
    from ecgbyte import tokenizer_utils as tu
    pct = _pct(gold)
    lst = str(tmp_path / "sampled.txt")
    with open(lst, "w") as f:
        for k, i in enumerate(gold["plf_order"]):
            f.write(("  " if k == 1 else "") + files[i] + ("   \n" if k % 2 else "\n"))
    for kw in ({}, {"batch": 2}, {"batch": 1}):
        assert tu.process_large_file(lst, pct, 2, **kw).encode() == gold["plf_all"].tobytes()
    assert tu.process_large_file(lst, pct, 3, n=4).encode() == gold["plf_n4"].tobytes()
    assert tu.process_large_file(lst, pct, 1, n=0) == ""
    for i in range(6):
        assert tu.process_ecg(files[i], pct).encode() == gold["pe_%d" % i].tobytes()


def test_normalize_all_integer_records_match_reference(gold):
    """An int16 .npy goes through the mirror on its raw integer values, as in the reference (no 1e-3 de-scaling), and
    both outputs of normalize_all agree with each other."""
    from ecgbyte import tokenizer_utils as tu
    pct = _pct(gold)
    r = gold["rec_3"]
    assert r.dtype == np.int16
    clipped, sym = tu.normalize_all(r, pct)
    assert "".join(sym.flatten()).encode() == gold["pe_3"].tobytes()
    want = np.minimum(np.floor(clipped * 26), 25).astype(np.uint8) + 97
    np.testing.assert_array_equal(np.frombuffer("".join(sym.flatten()).encode(), np.uint8), want.reshape(-1))


def test_analyze_token_distribution_matches_reference(gold, files):
    from ecgbyte import tokenizer_utils as tu
    with open(PKL, "rb") as f:
        vocab, merges = pickle.load(f)
    counts, lengths = tu.analyze_token_distribution([files[i] for i in gold["atd_files"]], merges, _pct(gold), num_workers=2)
    assert sorted(counts) == gold["atd_ids"].tolist()
    assert [counts[k] for k in sorted(counts)] == gold["atd_counts"].tolist()
    assert lengths == gold["atd_lengths"].tolist()


def test_expand_attention_device_path_matches_reference(gold):
    from ecgbyte import tokenizer_utils as tu
    with open(PKL, "rb") as f:
        vocab, merges = pickle.load(f)
    ids, att = gold["ea_ids_0"].tolist(), gold["ea_att_0"].tolist()
    got = tu.expand_attention(ids, att, vocab, merges=merges)      # device expands indices, values are the caller's
    assert got == gold["ea_out_0"].tolist()
    got = tu.expand_attention(gold["ea_ids_1"].tolist(), gold["ea_att_1"].tolist(), vocab, merges=merges)  # bytes > 127
    assert got == gold["ea_out_1"].tolist()
    assert tu.expand_attention(ids[:7], att[:4], vocab, merges=merges) == gold["ea_out_short"].tolist()


def test_train_tokenizer_call_sequence(tmp_path, oracle, gold, files):
    """train_tokenizer.py:19-45 and its self-check :47-64 with the drop-in modules in place of the reference's:
    percentiles dict from .npy, process_large_file, rust_bpe.byte_pair_encoding(text, num_merges, num_processes),
    save / load pickle, process_ecg, encode_text, decode_text, reverse_normalize_all."""
    import rust_bpe
    from ecgbyte.tokenizer_utils import (decode_text, encode_text, load_vocab_and_merges, process_ecg, process_large_file,
                                         reverse_normalize_all, save_vocab_and_merges)
    pct_path = str(tmp_path / "percentiles.npy")
    np.save(pct_path, {"percentile_1": gold["pct"][0], "percentile_99": gold["pct"][1]})
    percentiles = np.load(pct_path, allow_pickle=True).item()                       # :20
    lst = str(tmp_path / "sampled.txt")
    with open(lst, "w") as f:
        f.write("\n".join(files[i] for i in (0, 2, 5)) + "\n")
    num_processes, num_merges = 2, 120
    all_string_signals = process_large_file(lst, percentiles, num_processes)       # :25
    ids, vocab, merges = rust_bpe.byte_pair_encoding(all_string_signals, num_merges, num_processes)   # :29
    o_ids, o_vocab, o_merges = oracle.byte_pair_encoding(all_string_signals, num_merges, fast=True)
    assert ids == o_ids and vocab == o_vocab and merges == o_merges
    np.testing.assert_array_equal(np.array([m[0][-1] for m in merges]), np.array([m[0][-1] for m in o_merges]))
    name = str(tmp_path / ("tokenizer_%d.pkl" % num_merges))
    save_vocab_and_merges(vocab, merges, name)                                      # :38-39
    assert open(name, "rb").read() == open(PKL, "rb").read()                        # the reference's own pickle of the same table
    loaded_vocab, loaded_merges = load_vocab_and_merges(name)                       # :44
    new_ecg_signal = np.load(files[0])                                              # :47
    new_ecg_text = process_ecg(files[0], percentiles=percentiles)                   # :48
    encoded_ecg = encode_text(new_ecg_text, loaded_merges)                          # :53
    decoded_text = decode_text(encoded_ecg, loaded_vocab)                           # :58
    assert decoded_text == new_ecg_text                                             # :60
    decoded_signal = reverse_normalize_all(np.array(list(decoded_text)).reshape(new_ecg_signal.shape), percentiles)   # :62
    den = (percentiles["percentile_99"] + 0.5) - (percentiles["percentile_1"] - 0.5)
    inside = (new_ecg_signal > percentiles["percentile_1"] - 0.5) & (new_ecg_signal < percentiles["percentile_99"] + 0.5)
    assert np.max(np.abs(new_ecg_signal - decoded_signal)[inside]) <= den / 25.0     # :63 within one quantisation step


def test_track_encoding_matches_reference(gold):
    """tokenizer_utils.py:95-134 run by the reference itself: with the pickle's list-form merges nothing is merged;
    pair-form merges are applied in order with merge()'s greedy rule (here: the trainer's merge kernel), incl. (x,x)
    pairs on runs; segment_map is the byte range of every token."""
    from ecgbyte import tokenizer_utils as tu
    with open(PKL, "rb") as f:
        vocab, merges = pickle.load(f)
    txt = gold["te_text"].tobytes().decode()
    ids, seg = tu.track_encoding(txt, merges)
    assert ids == gold["te_ids_listform"].tolist() and [list(s) for s in seg] == gold["te_seg_listform"].tolist()
    pair_form = [((int(l), int(r)), 256 + i) for i, (l, r) in enumerate(gold["pairs"].tolist())]
    ids, seg = tu.track_encoding(txt, pair_form, verbose=False)
    assert ids == gold["te_ids_pairform"].tolist()
    assert [list(s) for s in seg] == gold["te_seg_pairform"].tolist()
    assert all(isinstance(s, tuple) for s in seg)
    ids, seg = tu.track_encoding("aaaaaaabaaaabbbbbbbbbaaa", [((97, 97), 300), ((98, 98), 301), ((300, 300), 302), ((301, 97), 303)])
    assert ids == gold["te_runs_ids"].tolist() and [list(s) for s in seg] == gold["te_runs_seg"].tolist()
    assert tu.track_encoding("", pair_form) == ([], [])
