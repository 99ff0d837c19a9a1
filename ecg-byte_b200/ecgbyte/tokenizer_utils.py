"""Mirror of the reference's ecg_byte/utils/tokenizer_utils.py for the hot path: same
names, arguments and return values, computed by libecgbyte.so.

  normalize_all, reverse_normalize_all        tokenizer_utils.py:14-28
  process_ecg, process_large_file             tokenizer_utils.py:56-59, 79-93
  encode_text, decode_text                    tokenizer_utils.py:71-77
  save/load_vocab_and_merges                  tokenizer_utils.py:62-69
  analyze_token_distribution                  tokenizer_utils.py:30-54
  expand_attention                            runners/interpret.py:106-111
"""
import pickle
from collections import Counter

import numpy as np
import torch

import rust_bpe
from .api import Quantizer, Vocab, token_histogram

ALPHABET = list("abcdefghijklmnopqrstuvwxyz")
_QUANT = {}


def _quantizer(percentiles, dtype):
    key = (float(percentiles["percentile_1"]), float(percentiles["percentile_99"]), dtype)
    q = _QUANT.get(key)
    if q is None:
        q = _QUANT[key] = Quantizer(percentiles, dtype=dtype)
    return q


def quantize_symbols(signal, percentiles):
    """signal (any shape; float64 / float32 / int16) -> uint8 symbol codes 'a'..'z'."""
    x = np.ascontiguousarray(signal)
    if x.dtype not in (np.float32, np.float64, np.int16):
        x = x.astype(np.float64)
    return _quantizer(percentiles, x.dtype).quantize_host(x)


def normalize_all(signal, percentiles):
    """-> (clipped_normalized float64 array, symbol_signal '<U1' array), tu.py:14-19."""
    sig = np.asarray(signal)
    codes = quantize_symbols(sig, percentiles)
    lo = percentiles["percentile_1"] - 0.5
    den = (percentiles["percentile_99"] + 0.5) - (percentiles["percentile_1"] - 0.5) + 1e-6
    x = torch.from_numpy(np.ascontiguousarray(sig, dtype=np.float64)).cuda()
    # tensor / tensor is an IEEE divide (tensor / python-scalar multiplies by a reciprocal)
    den_t = torch.tensor([float(den)], dtype=torch.float64, device=x.device)
    clipped = torch.clamp((x - float(lo)) / den_t, 0, 1).cpu().numpy()
    symbol_signal = codes.view("S1").astype("<U1").reshape(sig.shape)
    return clipped, symbol_signal


def reverse_normalize_all(symbol_signal, percentiles):
    """tu.py:22-28 (note: divides by len(ALPHABET) - 1 = 25, as the reference does)."""
    from .api import dequantize
    arr = np.asarray(symbol_signal)
    codes = np.ascontiguousarray(arr.astype("S1").view(np.uint8).reshape(arr.shape))
    return dequantize(torch.from_numpy(codes).cuda(), percentiles).cpu().numpy()


def process_ecg(ecg, percentiles):
    """path of a (C, L) .npy record -> its symbol string, lead-major (tu.py:56-59)."""
    sig = np.load(ecg) if isinstance(ecg, (str, bytes)) else np.asarray(ecg)
    return quantize_symbols(sig, percentiles).tobytes().decode("ascii")


def process_large_file(file_path, percentiles, num_processes=None, n=None):
    """tu.py:79-93: every listed record quantised and joined into ONE string, in file
    order.  `num_processes` is accepted for compatibility (the GPU path needs no pool)."""
    paths = []
    with open(file_path, "r") as f:
        for i, line in enumerate(f):
            if n is not None and i >= n:
                break
            paths.append(line.strip())
    parts = [process_ecg(p, percentiles) for p in paths]
    return "".join(parts)


def save_vocab_and_merges(vocab, merges, filename):
    with open(filename, "wb") as f:
        pickle.dump((vocab, merges), f)


def load_vocab_and_merges(filename):
    with open(filename, "rb") as f:
        vocab, merges = pickle.load(f)
    return vocab, merges


def encode_text(text, merges):
    return rust_bpe.encode_text(text, merges)


def decode_text(encoded_ids, vocab):
    return "".join(vocab[i] for i in encoded_ids)


def analyze_token_distribution(test_data, merges, percentiles, num_workers=None, batch=256):
    """(token_counts: Counter, token_lengths: list[int]) over the records whose .npy paths are listed in
    test_data (tu.py:30-54).  Records are quantised + encoded in batches on the GPU and counted there;
    num_workers is accepted and ignored (the reference fans the files out over a process pool)."""
    vocab = Vocab(merges)
    n_ids = max([255] + [int(i) for _, i in merges]) + 1
    counts = torch.zeros((n_ids,), dtype=torch.int64, device="cuda")
    token_lengths = []
    paths = list(test_data)
    for s in range(0, len(paths), batch):
        recs = [np.load(p) for p in paths[s:s + batch]]
        shapes = {r.shape for r in recs}
        groups = [recs] if len(shapes) == 1 else [[r] for r in recs]  # ragged shapes: one record per call
        for g in groups:
            x = np.ascontiguousarray(np.stack(g))
            if x.dtype not in (np.float32, np.float64, np.int16):
                x = x.astype(np.float64)
            q = _quantizer(percentiles, x.dtype)
            tokens, lens = vocab.encode_batch(q, torch.from_numpy(x).cuda())
            token_histogram(tokens, lens, n_ids, counts)
            token_lengths.extend(int(v) for v in lens.cpu().tolist())
    c = counts.cpu().numpy()
    return Counter({int(i): int(c[i]) for i in np.nonzero(c)[0]}), token_lengths


def expand_attention(encoded_ids, attention_sequence, vocab, merges=None):
    """runners/interpret.py:106-111: each token's attention value repeated once per symbol of the token.
    With `merges` the expansion runs on the GPU (lengths from the vocabulary's decode table); without, the
    lengths come from the vocab strings on the host exactly as the reference does."""
    if merges is None:
        out = []
        for i, a in zip(encoded_ids, attention_sequence):
            out.extend([a] * len(vocab[i]))
        return out
    n = min(len(encoded_ids), len(attention_sequence))
    if n == 0:
        return []
    tok = torch.tensor(list(encoded_ids)[:n], dtype=torch.int32, device="cuda").view(1, n)
    att = torch.tensor(list(attention_sequence)[:n], dtype=torch.float32, device="cuda").view(1, n)
    lens = torch.tensor([n], dtype=torch.int32, device="cuda")
    total = sum(len(vocab[int(i)]) for i in list(encoded_ids)[:n])
    out, out_len = Vocab(merges).expand_attention(tok, lens, att, max(total, 1))
    return out[0, : int(out_len[0])].cpu().tolist()
