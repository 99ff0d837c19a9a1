"""Pure-Python restatement of the reference tokenizer (small cases only).

TEST INFRASTRUCTURE: a second, independent statement of the same semantics as
oracle/ecgb_oracle.c, written from the reference text, used to cross-check the C
oracle.  Plain loops / dicts, no shared code with the C file.

Reference (paths under /root/reference/ecg_byte):
  rust_bpe/src/lib.rs:10-26   merge
  rust_bpe/src/lib.rs:28-48   get_stats
  rust_bpe/src/lib.rs:58-125  byte_pair_encoding
  rust_bpe/src/lib.rs:127-193 TrieNode / encode_text
  utils/tokenizer_utils.py:14-19 normalize_all ; :75-77 decode_text

Tie rule (the reference's is hash-order dependent): maximum
count, then lexicographically smallest (left, right).
"""
import numpy as np

ALPHABET = list("abcdefghijklmnopqrstuvwxyz")


def normalize_all_symbols(signal, p1, p99):
    """tu.py:14-19 in float64; returns the symbol codes (uint8 'a'..'z')."""
    s = np.asarray(signal).astype(np.float64)
    normalized = (s - (p1 - 0.5)) / ((p99 + 0.5) - (p1 - 0.5) + 1e-6)
    clipped = np.clip(normalized, 0, 1)
    with np.errstate(invalid="ignore"):
        q = np.minimum(np.floor(clipped * len(ALPHABET)), len(ALPHABET) - 1)
        q = np.where(np.isnan(q), 0, q).astype(np.uint8)
    return (q + 97).astype(np.uint8)


def merge(ids, pair, new_id):
    out = []
    i = 0
    n = len(ids)
    while i < n:
        if i + 1 < n and (ids[i], ids[i + 1]) == pair:
            out.append(new_id)
            i += 2
        else:
            out.append(ids[i])
            i += 1
    return out


def get_stats(ids):
    acc = {}
    for a, b in zip(ids, ids[1:]):
        acc[(a, b)] = acc.get((a, b), 0) + 1
    return acc


def byte_to_string(b):
    return chr(b) if b <= 127 else "<%d>" % b


def byte_pair_encoding(text, num_merges, num_threads=1, tie_log=None):
    data = text.encode("utf-8") if isinstance(text, str) else bytes(text)
    ids = list(data)
    vocab = {i: byte_to_string(i) for i in range(256)}
    vocab_tokens = {i: [i] for i in range(256)}
    merges = []
    for i in range(num_merges):
        pairs = get_stats(ids)
        if not pairs:
            break
        top = max(pairs.values())
        tied = sorted(p for p, c in pairs.items() if c == top)
        best = tied[0]
        if tie_log is not None:
            tie_log.append((i, top, len(tied)))
        new_id = 256 + i
        ids = merge(ids, best, new_id)
        vocab[new_id] = vocab[best[0]] + vocab[best[1]]
        vocab_tokens[new_id] = vocab_tokens[best[0]] + vocab_tokens[best[1]]
        merges.append((list(vocab_tokens[new_id]), new_id))
    return ids, vocab, merges


class _Node:
    __slots__ = ("children", "token_id")

    def __init__(self):
        self.children = {}
        self.token_id = None


def encode_text(text, merges):
    data = text.encode("utf-8") if isinstance(text, str) else bytes(text)
    ids = list(data)
    root = _Node()

    def insert(seq, tid):
        node = root
        for s in seq:
            node = node.children.setdefault(s, _Node())
        node.token_id = tid

    for b in range(256):
        insert([b], b)
    for seq, tid in merges:
        insert(seq, tid)
    out = []
    i = 0
    n = len(ids)
    while i < n:
        node = root
        match_len, match_id = 0, None
        for j in range(i, n):
            child = node.children.get(ids[j])
            if child is None:
                break
            node = child
            if node.token_id is not None:
                match_len, match_id = j - i + 1, node.token_id
        if match_id is not None:
            out.append(match_id)
            i += match_len
        else:
            out.append(ids[i])
            i += 1
    return out


def decode_text(encoded_ids, vocab):
    return "".join(vocab[i] for i in encoded_ids)


# ---------------------------------------------------------------------------
# rows either side of the encoder
# ---------------------------------------------------------------------------
def reverse_normalize_all(symbol_codes, p1, p99):
    """tokenizer_utils.py:22-28 on uint8 symbol codes ('a'..'z')."""
    min_vals = p1 - 0.5
    max_vals = p99 + 0.5
    scaled = np.asarray(symbol_codes).astype(np.int64) - 97
    clipped = scaled / (len(ALPHABET) - 1)
    return clipped * (max_vals - min_vals) + min_vals


def decode_symbols(tokens, merges):
    """decode_text (tokenizer_utils.py:75-77) as bytes: concatenation of each token's sequence."""
    table = {i: [i] for i in range(256)}
    for seq, tid in merges:
        table[tid] = list(seq)
    out = []
    for t in tokens:
        out.extend(table[int(t)])
    return np.array(out, np.uint8)


def prepare_training(signal_ids, question, answer, pad_to_max, pad_id, bos_id, eos_id, sig_start_id, sig_end_id):
    """data_loader.py:26-31 and 101-132 (ECGTokenDataset._prepare_training) in plain Python.
    signal_ids are already LLM ids (data_loader.py:80).  Returns (input_ids, attn_mask, labels, position_ids)."""
    sig = list(signal_ids)
    qa_len = len(question) + len(answer)
    available = pad_to_max - qa_len
    if available < 0:
        raise AssertionError("question + answer longer than pad_to_max")
    if len(sig) > available:
        sig = [bos_id, sig_start_id] + sig[:available] + [sig_end_id]
    elif len(sig) < available:
        sig = [pad_id] * (available - len(sig)) + [bos_id, sig_start_id] + sig + [sig_end_id]
    else:
        sig = [bos_id, sig_start_id] + sig + [sig_end_id]
    sample = sig + list(question) + list(answer) + [eos_id]
    labels = [-100] * (len(sig) + len(question)) + list(answer) + [eos_id]
    mask = [0 if t == pad_id else 1 for t in sample]
    pos, run = [], 0
    for m in mask:
        run += m
        pos.append(run - 1 if m else 0)
    assert len(sample) == pad_to_max + 4
    return (np.array(sample, np.int64), np.array(mask, np.float32), np.array(labels, np.int64), np.array(pos, np.int64))


def expand_attention(encoded_ids, attention_sequence, vocab):
    """runners/interpret.py:106-111."""
    out = []
    for i, a in zip(encoded_ids, attention_sequence):
        out.extend([a] * len(vocab[i]))
    return out


def token_distribution(encoded_records):
    """tokenizer_utils.py:30-54 without the file I/O: Counter over all ids + the per-record lengths."""
    from collections import Counter
    counts, lengths = Counter(), []
    for ids in encoded_records:
        counts.update(int(i) for i in ids)
        lengths.append(len(ids))
    return counts, lengths


def normalize_all_as_written(signal, p1, p99):
    """tu.py:14-19 the way the reference EXECUTES it: float arithmetic in NumPy, then one Python call per sample
    through np.vectorize to map codes to letters, and (tu.py:59 / data_loader.py:75) ''.join over the flattened
    '<U1' array.  Only used to time the reference's own NumPy path beside the C port (bench.py cpu_baseline)."""
    s = np.asarray(signal)
    normalized = (s - (p1 - 0.5)) / ((p99 + 0.5) - (p1 - 0.5) + 1e-6)
    clipped = np.clip(normalized, 0, 1)
    scaled = np.minimum(np.floor(clipped * len(ALPHABET)), len(ALPHABET) - 1).astype(np.uint8)
    symbols = np.vectorize(lambda c: ALPHABET[c])(scaled)
    return clipped, "".join(symbols.flatten())
