"""Generates tests/golden/post_reference.npz by running the REFERENCE's own Python code for
the rows either side of the encoder (build container only; needs /root/reference):
  * ecg_byte.utils.tokenizer_utils.decode_text / reverse_normalize_all   (tu.py:22-28, 75-77)
  * ecg_byte.data_loader.ECGTokenDataset._prepare_training                (data_loader.py:101-132)
    called unbound on a stub `self` that carries only the ids / pad_to_max it reads.
Run:  python oracle/make_golden_post.py"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from make_golden import import_reference_tu  # noqa: E402


def main():
    tu = import_reference_tu()
    import ecg_byte.data_loader as dl
    from ecgbyte import synth
    import oracle as O

    rng = np.random.default_rng(77)
    store = {}
    # ---- decode path: records -> symbols -> tokens (oracle) -> reference decode_text / reverse_normalize_all
    x = synth.corpus(11, 3, L=400, dtype=np.float64)
    pct = synth.percentiles(x, seed=11)
    sym = O.quantize(x, pct["percentile_1"], pct["percentile_99"])
    ids, vocab, merges = O.byte_pair_encoding(sym.reshape(-1).tobytes().decode(), 150, fast=True)
    store["dec_pairs"] = np.array([[0, 0]], np.uint32)  # placeholder, replaced below
    _, pairs, _, _ = O.train_pairs(sym.reshape(-1), 150, fast=True)
    store["dec_pairs"] = pairs
    store["dec_pct"] = np.array([pct["percentile_1"], pct["percentile_99"]])
    for r in range(3):
        s = sym[r].reshape(-1).tobytes().decode()
        enc = O.encode_text(s, merges)
        text = tu.decode_text(enc, vocab)                                   # the reference
        assert text == s
        vals = tu.reverse_normalize_all(np.array(list(text)).reshape(x[r].shape), pct)   # the reference
        store["dec_tokens_%d" % r] = np.array(enc, np.int32)
        store["dec_text_%d" % r] = np.frombuffer(text.encode(), np.uint8)
        store["dec_values_%d" % r] = vals
    # ---- ECGTokenDataset._prepare_training
    cases = []
    for pad_to_max, nsig, nq, na in [(1020, 4600, 12, 30), (1020, 300, 9, 5), (64, 40, 10, 14), (64, 50, 10, 14),
                                     (64, 39, 10, 14), (32, 0, 3, 4), (40, 7, 20, 20), (16, 100, 0, 1)]:
        stub = types.SimpleNamespace(args=types.SimpleNamespace(pad_to_max=pad_to_max), pad_id=128256, bos_id=128000,
                                     eos_id=128001, sig_start_id=[133515], sig_end_id=[133516])
        sig = rng.integers(128257, 133513, size=nsig).tolist()
        q = rng.integers(0, 128000, size=nq).tolist()
        a = rng.integers(0, 128000, size=na).tolist()
        out = dl.ECGTokenDataset._prepare_training(stub, list(sig), list(q), list(a), None, None, None)   # the reference
        k = len(cases)
        store["pack_cfg_%d" % k] = np.array([pad_to_max, 128256, 128000, 128001, 133515, 133516], np.int64)
        store["pack_sig_%d" % k] = np.array(sig, np.int64)
        store["pack_q_%d" % k] = np.array(q, np.int64)
        store["pack_a_%d" % k] = np.array(a, np.int64)
        store["pack_ids_%d" % k] = out["tokenized_signal"].numpy()
        store["pack_attn_%d" % k] = out["attn_mask"].numpy()
        store["pack_labels_%d" % k] = out["quantized_signal_ids_input"].numpy()
        store["pack_pos_%d" % k] = out["position_ids"].numpy()
        cases.append(k)
    store["pack_n"] = np.array([len(cases)])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "post_reference.npz"), **store)
    print("wrote post_reference.npz:", len(cases), "pack cases, 3 decode records")


if __name__ == "__main__":
    main()
