"""Drop-in for the reference's PyO3 module `rust_bpe`
(/root/reference/ecg_byte/rust_bpe/src/lib.rs:195-199), backed by libecgbyte.so.

    byte_pair_encoding(text, num_merges, num_threads) -> (ids, vocab, merges)   lib.rs:58-125
    encode_text(text, merges) -> ids                                            lib.rs:149-193

Argument and return types are the reference's: `ids: list[int]`,
`vocab: dict[int, str]`, `merges: list[tuple[list[int], int]]`, so pickles written by
either implementation load in the other (tokenizer_utils.py:62-69).

Differences (stated): equal-count ties are broken by the smallest
(left, right) pair instead of hash-map order; `num_threads` is accepted and ignored;
no progress bar / timing print; the vocab dict is in ascending id order.
"""
import numpy as _np

_VOCAB_CACHE = {}
_CACHE_MAX = 4


def _byte_to_string(b):
    # lib.rs:50-56
    return chr(b) if b <= 127 else "<%d>" % b


def _check_text(text):
    if not isinstance(text, str):
        raise TypeError("argument 'text': 'str' expected, got %s" % type(text).__name__)
    return text.encode("utf-8")


def byte_pair_encoding(text, num_merges, num_threads=0):
    from ecgbyte.api import Trainer, expand_merges

    data = _check_text(text)
    if not isinstance(num_merges, int) or num_merges < 0:
        raise TypeError("argument 'num_merges': non-negative int expected")
    if not isinstance(num_threads, int) or num_threads < 0:
        raise TypeError("argument 'num_threads': non-negative int expected")
    n = len(data)
    tr = Trainer(max(n, 1), num_merges)
    tr.load(data)
    pairs, _, _ = tr.run(num_merges)
    ids = tr.ids().tolist()
    seq, off = expand_merges(pairs)
    vocab = {i: _byte_to_string(i) for i in range(256)}
    merges = []
    seq_l = seq.tolist()
    off_l = off.tolist()
    for i, (l, r) in enumerate(pairs.tolist()):
        new_id = 256 + i
        vocab[new_id] = vocab[l] + vocab[r]                      # lib.rs:101-104
        merges.append((seq_l[off_l[i]: off_l[i + 1]], new_id))   # lib.rs:106-110
    return ids, vocab, merges


def _fingerprint(merges):
    """Cheap content check of a merges list: its length and a spread of 16 entries (sequence length, id, first and
    last symbol).  Enough to notice a list that was edited in place; an edit that keeps all sampled values is not
    detected -- pass a new list object (the reference rebuilds its trie on every call, lib.rs:153-161)."""
    n = len(merges)
    fp = [n]
    for k in range(16):
        if n == 0:
            break
        seq, tid = merges[(k * n) // 16]
        fp.append((len(seq), tid, seq[0] if len(seq) else -1, seq[-1] if len(seq) else -1))
    return tuple(fp)


def _vocab_for(merges):
    """The reference rebuilds the trie on every call (lib.rs:153-161); here the flattened device trie is cached
    per merges object: identity + a content fingerprint (see _fingerprint), and per device."""
    import torch
    from ecgbyte.api import Vocab

    for item in merges[:1]:
        if not (isinstance(item, (tuple, list)) and len(item) == 2):
            raise TypeError("argument 'merges': expected a list of (sequence, id) tuples")
    key = (id(merges), torch.cuda.current_device() if torch.cuda.is_available() else -1)
    fp = _fingerprint(merges)
    hit = _VOCAB_CACHE.get(key)
    if hit is not None and hit[0] is merges and hit[2] == fp:
        return hit[1]
    for item in merges:
        if not (isinstance(item, (tuple, list)) and len(item) == 2):
            raise TypeError("argument 'merges': expected a list of (sequence, id) tuples")
    v = Vocab(merges=merges)
    _VOCAB_CACHE.pop(key, None)
    if len(_VOCAB_CACHE) >= _CACHE_MAX:
        _VOCAB_CACHE.pop(next(iter(_VOCAB_CACHE)))
    _VOCAB_CACHE[key] = (merges, v, fp)
    return v


def encode_text(text, merges):
    data = _check_text(text)
    if not isinstance(merges, list):
        raise TypeError("argument 'merges': 'list' expected, got %s" % type(merges).__name__)
    return _vocab_for(merges).encode_text(data)


__all__ = ["byte_pair_encoding", "encode_text"]
