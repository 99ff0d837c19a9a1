"""Generates tests/golden/callsites_reference.npz and tests/golden/ref_vocab_merges.pkl by running the REFERENCE's own
Python call sites of the hot path (build container only; needs /root/reference):
  * ecg_byte.utils.tokenizer_utils.process_large_file                    (tu.py:79-93: file order, .strip(), n cap)
  * ecg_byte.utils.tokenizer_utils.analyze_token_distribution            (tu.py:30-54)
  * ecg_byte.runners.interpret.expand_attention                          (runners/interpret.py:106-111)
  * ecg_byte.utils.tokenizer_utils.save_vocab_and_merges                 (tu.py:62-64)
  * ecg_byte.utils.tokenizer_utils.track_encoding                        (tu.py:95-134)
The native module rust_bpe cannot be built here (no Rust toolchain); where a reference function calls
rust_bpe.encode_text the stub installed in sys.modules forwards to the C restatement of lib.rs:149-193
(oracle/ecgb_oracle.c) -- everything around that call is the reference's unmodified code.
expand_attention is taken from the reference file by its AST (the module's imports need packages that are absent here).
Run:  python oracle/make_golden_callsites.py"""
import ast
import os
import pickle
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200"), os.path.join(ROOT, "oracle")]
from make_golden import REF, import_reference_tu  # noqa: E402


def reference_function(path, name):
    """the function `name` of the reference file `path`, compiled from its own source text"""
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {}
            exec(compile(mod, path, "exec"), ns)
            return ns[name]
    raise KeyError(name)


def main():
    import oracle as O
    from ecgbyte import synth
    tu = import_reference_tu()
    sys.modules["rust_bpe"].encode_text = lambda text, merges: O.encode_text(text, merges)
    store = {}

    # ---- records on disk
    x = synth.corpus(21, 5, L=250, dtype=np.float64)
    pct = synth.percentiles(x, seed=21)
    recs = [x[0], x[1].astype(np.float32), x[2], np.clip(np.round(x[3] * 1000.0), -32768, 32767).astype(np.int16),
            x[4][:3, :100].copy(), x[4]]
    store["pct"] = np.array([pct["percentile_1"], pct["percentile_99"]])
    for i, r in enumerate(recs):
        store["rec_%d" % i] = r
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for i, r in enumerate(recs):
            p = os.path.join(d, "ecg_%d.npy" % i)
            np.save(p, r)
            paths.append(p)
        order = [2, 0, 5, 3, 1, 4]
        lst = os.path.join(d, "sampled.txt")
        with open(lst, "w") as f:
            for k, i in enumerate(order):
                f.write(("  " if k == 1 else "") + paths[i] + ("   \n" if k % 2 else "\n"))   # stray blanks: .strip()
        store["plf_order"] = np.array(order)
        store["plf_all"] = np.frombuffer(tu.process_large_file(lst, pct, 2).encode(), np.uint8)          # the reference
        store["plf_n4"] = np.frombuffer(tu.process_large_file(lst, pct, 2, n=4).encode(), np.uint8)      # the reference
        for i in range(len(recs)):
            store["pe_%d" % i] = np.frombuffer(tu.process_ecg(paths[i], pct).encode(), np.uint8)         # the reference

        # ---- a merge table for the analytics
        corpus = "".join(tu.process_ecg(paths[i], pct) for i in (0, 2, 5))
        ids, vocab, merges = O.byte_pair_encoding(corpus, 120, fast=True)
        _, pairs, _, _ = O.train_pairs(np.frombuffer(corpus.encode(), np.uint8), 120, fast=True)
        store["pairs"] = pairs
        counts, lengths = tu.analyze_token_distribution([paths[i] for i in (0, 1, 2, 5, 4)], merges, pct, num_workers=2)  # the reference
        store["atd_files"] = np.array([0, 1, 2, 5, 4])
        store["atd_ids"] = np.array(sorted(counts), np.int64)
        store["atd_counts"] = np.array([counts[k] for k in sorted(counts)], np.int64)
        store["atd_lengths"] = np.array(lengths, np.int64)

    # ---- expand_attention
    ea = reference_function(os.path.join(REF, "ecg_byte", "runners", "interpret.py"), "expand_attention")
    rng = np.random.default_rng(5)
    enc = O.encode_text(corpus[:3000], merges)
    att = rng.random(len(enc)).tolist()
    store["ea_ids_0"] = np.array(enc, np.int64)
    store["ea_att_0"] = np.array(att, np.float64)
    store["ea_out_0"] = np.array(ea(enc, att, vocab), np.float64)                                         # the reference
    enc1 = [97, 200, 256, 255, 98]          # raw bytes > 127: the vocab string "<200>" has 5 characters
    att1 = [0.1, 0.2, 0.3, 0.4, 0.5]
    store["ea_ids_1"] = np.array(enc1, np.int64)
    store["ea_att_1"] = np.array(att1, np.float64)
    store["ea_out_1"] = np.array(ea(enc1, att1, vocab), np.float64)                                       # the reference
    store["ea_out_short"] = np.array(ea(enc[:7], att[:4], vocab), np.float64)                              # zip stops at the shorter

    # ---- track_encoding: (a) the pickle's own format -- the expanded sequence is a list, so `(a, b) == pair` never holds
    #      and nothing is merged; (b) pair-form merges, which it does apply, incl. an (x,x) pair on a run
    txt = corpus[:1500]
    ids_a, seg_a = tu.track_encoding(txt, merges, verbose=False)                                           # the reference
    store["te_text"] = np.frombuffer(txt.encode(), np.uint8)
    store["te_ids_listform"] = np.array(ids_a, np.int64)
    store["te_seg_listform"] = np.array(seg_a, np.int64)
    pair_form = [((int(l), int(r)), 256 + i) for i, (l, r) in enumerate(pairs.tolist())]
    ids_b, seg_b = tu.track_encoding(txt, pair_form, verbose=False)                                        # the reference
    store["te_ids_pairform"] = np.array(ids_b, np.int64)
    store["te_seg_pairform"] = np.array(seg_b, np.int64)
    runs = "aaaaaaabaaaabbbbbbbbbaaa"
    pf2 = [((97, 97), 300), ((98, 98), 301), ((300, 300), 302), ((301, 97), 303)]
    ids_c, seg_c = tu.track_encoding(runs, pf2)                                                            # the reference
    store["te_runs_ids"] = np.array(ids_c, np.int64)
    store["te_runs_seg"] = np.array(seg_c, np.int64)

    out = os.path.join(ROOT, "tests", "golden")
    np.savez_compressed(os.path.join(out, "callsites_reference.npz"), **store)
    tu.save_vocab_and_merges(vocab, merges, os.path.join(out, "ref_vocab_merges.pkl"))                     # the reference
    v2, m2 = tu.load_vocab_and_merges(os.path.join(out, "ref_vocab_merges.pkl"))
    assert v2 == vocab and m2 == merges
    print("wrote callsites_reference.npz (%d arrays) and ref_vocab_merges.pkl" % len(store))


if __name__ == "__main__":
    main()
