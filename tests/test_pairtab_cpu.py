"""CPU suite for the flattened trie the fused encoder walks (csrc/pairtab.cu, csrc/trie_host.h):
the table built by libecgbyte.so (host code, no device) is walked by an executable model of the
kernel's rules (tests/pairtab_model.py) and compared with the oracle's trie encoder
(lib.rs:163-190 restated in oracle/ecgb_oracle.c)."""
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from pairtab_model import PairTable

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ptbxl_1000_m5000.npz")


def _table(oracle, pairs):
    seq, off = oracle.expand(pairs)
    ids = np.arange(256, 256 + len(pairs), dtype=np.uint32)
    return PairTable(seq, off, ids), oracle.Trie(flat=(seq, off, ids))


def test_layout_invariants(oracle, small_table):
    pairs, _, _ = small_table
    pt, _ = _table(oracle, pairs)
    assert pt.W == 5 and pt.NC == 26 and pt.SM == 26 and pt.SE == 27
    e = pt.ent
    live = e != 0xFFFFFFFF
    assert live.sum() == pt.n_used
    assert np.all((e[live] >> 14) & 3 == 0)                       # bits 14-15 are zero: base << 2 == e >> 14
    assert len(e) >= max(pt.dead_base, int(np.flatnonzero(live).max()) + 1) + (1 << (2 * pt.W))
    # every slot lies at base + code of exactly one row: code stored == slot - base for the owning row,
    # so no probe from the dead base (or from a foreign row) can match
    codes = (e >> 2) & 0xFFF
    slots = np.flatnonzero(live)
    bases = slots - codes[live]
    assert np.all(bases >= 0) and pt.dead_base not in set(bases.tolist())
    for x in range(1 << (2 * pt.W)):
        ee = int(e[pt.dead_base + x])
        assert ((ee >> 2) & 0xFFF) != x


def test_model_equals_oracle_on_ecg_records(oracle, small_corpus, small_table):
    x, pct = small_corpus
    pairs, _, _ = small_table
    pt, trie = _table(oracle, pairs)
    sym = oracle.quantize(x, pct["percentile_1"], pct["percentile_99"]).reshape(x.shape[0], -1)
    for r in range(4):
        np.testing.assert_array_equal(pt.encode(sym[r].tobytes()), trie.encode(sym[r]))


def test_model_equals_oracle_on_the_5000_merge_table(oracle):
    from ecgbyte import synth
    f = np.load(FIXTURE)
    pt, trie = _table(oracle, f["pairs"].astype(np.uint32))
    assert pt.ent.size * 4 < 64 * 1024, "the config-2 table is meant to stay well inside shared memory"
    x = synth.corpus(99, 2, 5000, np.float32)
    sym = oracle.quantize(x, f["pct"][0], f["pct"][1]).reshape(2, -1)
    for r in range(2):
        np.testing.assert_array_equal(pt.encode(sym[r].tobytes()), trie.encode(sym[r]))


def test_known_answers(oracle):
    # lib.rs semantics: longest match, not rank order; a later duplicate wins;
    # interior nodes that are no tokens; bytes outside every merge are their own tokens
    merges = [([98, 99], 256), ([97, 98], 257)]
    seq, off, ids = oracle.flatten_merges(merges)
    pt = PairTable(seq, off, ids)
    assert pt.encode(b"abc").tolist() == [257, 99]
    merges = [([97, 97, 97, 97], 300), ([97, 97], 301), ([97, 97], 302)]
    seq, off, ids = oracle.flatten_merges(merges)
    pt = PairTable(seq, off, ids)
    assert pt.encode(b"aaaaaaa").tolist() == [300, 302, 97]      # aaaa | aa | a ; aaa is no token
    assert pt.encode(b"aaXaaa!").tolist() == [302, 88, 302, 97, 33]
    assert pt.encode(b"").tolist() == []


@settings(max_examples=60, deadline=None)
@given(st.data())
def test_model_equals_oracle_random_tables(oracle, data):
    # small alphabets make deep, branchy tries with many non-terminal interior nodes
    alpha = data.draw(st.sampled_from([b"ab", b"abc", b"abcz", b"aZ!q"]))
    n_merges = data.draw(st.integers(0, 40))
    merges = []
    for i in range(n_merges):
        ln = data.draw(st.integers(2, 9))
        s = [alpha[data.draw(st.integers(0, len(alpha) - 1))] for _ in range(ln)]
        merges.append((s, 256 + data.draw(st.integers(0, 500))))
    seq, off, ids = oracle.flatten_merges(merges)
    pt = PairTable(seq, off, ids)
    text = bytes(data.draw(st.lists(st.sampled_from(list(alpha) + [0x5F]), min_size=0, max_size=120)))
    assert pt.encode(text).tolist() == oracle.encode_text(text, merges)
