"""Mirror of the reference's ecg_byte/utils/tokenizer_utils.py for the hot path: same
names, arguments and return values, computed by libecgbyte.so.

  normalize_all, reverse_normalize_all        tokenizer_utils.py:14-28
  process_ecg, process_large_file             tokenizer_utils.py:56-59, 79-93
  encode_text, decode_text                    tokenizer_utils.py:71-77
  save/load_vocab_and_merges                  tokenizer_utils.py:62-69
  analyze_token_distribution                  tokenizer_utils.py:30-54
  expand_attention                            runners/interpret.py:106-111
  track_encoding                              tokenizer_utils.py:95-134
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
    key = (float(percentiles["percentile_1"]), float(percentiles["percentile_99"]), dtype, torch.cuda.current_device())
    q = _QUANT.get(key)
    if q is None:
        q = _QUANT[key] = Quantizer(percentiles, dtype=dtype)
    return q


def _as_samples(signal):
    """The reference computes (signal - lo) / den on whatever np.load returned: float64 / float32 values as they
    are, every integer type (int16 included) on its raw integer values.  Only the two float types have a kernel of
    their own here; everything else is widened to float64 first, exactly like NumPy's promotion does.  (The 1e-3
    de-scaling int16 path exists only behind the explicit Quantizer(dtype=int16, i16_scale=...) API.)"""
    x = np.ascontiguousarray(signal)
    if x.dtype not in (np.float32, np.float64):
        x = x.astype(np.float64)
    return x


def quantize_symbols(signal, percentiles):
    """signal (any shape, any numeric dtype) -> uint8 symbol codes 'a'..'z'."""
    x = _as_samples(signal)
    return _quantizer(percentiles, x.dtype).quantize_host(x)


def normalize_all(signal, percentiles):
    """-> (clipped_normalized array, symbol_signal '<U1' array), tu.py:14-19.  The symbols (the only output the
    hot path uses: data_loader.py:74, tu.py:58) come from the device; the clipped values, which no caller on the
    path reads, are the reference's own NumPy expression."""
    sig = np.asarray(signal)
    codes = quantize_symbols(sig, percentiles)
    normalized = (sig - (percentiles["percentile_1"] - 0.5)) / (
        (percentiles["percentile_99"] + 0.5) - (percentiles["percentile_1"] - 0.5) + 1e-6)
    clipped = np.clip(normalized, 0, 1)
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


def process_large_file(file_path, percentiles, num_processes=None, n=None, batch=256):
    """tu.py:79-93: every listed record quantised and joined into ONE string, in file order; `n` caps the number
    of lines read, each line is .strip()ped.  The reference fans the files out over `num_processes` worker
    processes; here that many threads read the .npy files and runs of equally shaped records go to the device
    as one batch (one copy in, one kernel, one copy out per up to `batch` files)."""
    from concurrent.futures import ThreadPoolExecutor
    paths = []
    with open(file_path, "r") as f:
        for i, line in enumerate(f):
            if n is not None and i >= n:
                break
            paths.append(line.strip())
    parts = []
    with ThreadPoolExecutor(max_workers=max(1, int(num_processes or 1))) as pool:
        for s0 in range(0, len(paths), batch):
            recs = [_as_samples(r) for r in pool.map(np.load, paths[s0:s0 + batch])]
            k = 0
            while k < len(recs):  # a run of records with one shape and dtype
                e = k + 1
                while e < len(recs) and recs[e].shape == recs[k].shape and recs[e].dtype == recs[k].dtype:
                    e += 1
                codes = _quantizer(percentiles, recs[k].dtype).quantize_host(np.stack(recs[k:e]))
                parts.append(codes.tobytes().decode("ascii"))
                k = e
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
    vocab = rust_bpe._vocab_for(merges)
    n_ids = max([255] + [int(i) for _, i in merges]) + 1
    counts = torch.zeros((n_ids,), dtype=torch.int64, device="cuda")
    token_lengths = []
    paths = list(test_data)
    for s in range(0, len(paths), batch):
        recs = [np.load(p) for p in paths[s:s + batch]]
        shapes = {r.shape for r in recs}
        groups = [recs] if len(shapes) == 1 else [[r] for r in recs]  # ragged shapes: one record per call
        for g in groups:
            x = _as_samples(np.stack(g))
            q = _quantizer(percentiles, x.dtype)
            tokens, lens = vocab.encode_batch(q, torch.from_numpy(x).cuda())
            token_histogram(tokens, lens, n_ids, counts)
            token_lengths.extend(int(v) for v in lens.cpu().tolist())
    c = counts.cpu().numpy()
    return Counter({int(i): int(c[i]) for i in np.nonzero(c)[0]}), token_lengths


def expand_attention(encoded_ids, attention_sequence, vocab, merges=None):
    """runners/interpret.py:106-111: each token's attention value repeated len(vocab[id]) times.  Without `merges`
    that is done on the host exactly as the reference does.  With `merges` the device expands token INDICES (one per
    base symbol, from the vocabulary's decode table) and the result is gathered from the caller's own values, so the
    returned objects are the originals, not float32 roundings.  A vocab string is longer than its symbol count only
    for raw bytes > 127 (spelled "<200>", lib.rs:50-56); such input takes the host path."""
    ids = list(encoded_ids)
    att = list(attention_sequence)
    n = min(len(ids), len(att))

    def host():
        out = []
        for i, a in zip(ids, att):
            out.extend([a] * len(vocab[i]))
        return out

    if merges is None or n == 0 or any(127 < int(i) < 256 for i in ids[:n]):
        return host()
    v = rust_bpe._vocab_for(merges)
    if getattr(v, "_has_high_bytes", None) is None:
        v._has_high_bytes = bool((v.flat[0] > 127).any())
    if v._has_high_bytes or n >= (1 << 24):
        return host()
    tok = torch.tensor(ids[:n], dtype=torch.int32, device="cuda").view(1, n)
    idx = torch.arange(n, dtype=torch.float32, device="cuda").view(1, n)  # exact below 2^24
    lens = torch.tensor([n], dtype=torch.int32, device="cuda")
    cap = n * int(v.info()["max_token_len"])
    out, out_len = v.expand_attention(tok, lens, idx, max(cap, 1))
    which = out[0, : int(out_len[0])].to(torch.int64).cpu().tolist()
    return [att[j] for j in which]


def track_encoding(text, merges, verbose=True):
    """tokenizer_utils.py:95-134: apply `merges` to the UTF-8 bytes of `text` in order, each with merge()'s greedy
    left-to-right rule (lib.rs:10-26), and return (ids, segment_map) where segment_map[k] = (start, end) byte range of
    token k.  The reference tests `(ids[i], ids[i + 1]) == pair`, which can only hold when `pair` is a 2-tuple: with the
    pickle's own format (`pair` = the expanded sequence, a list) nothing is ever merged -- same here.  Pair-form merges
    run through the trainer's merge kernel on the device; `verbose` only switched a progress bar."""
    from .api import Trainer
    data = text.encode("utf-8")
    todo = [(p, nid) for p, nid in merges
            if isinstance(p, tuple) and len(p) == 2 and all(isinstance(x, int) and not isinstance(x, bool) for x in p)]
    n = len(data)
    if not todo or n == 0:
        return list(data), [(i, i + 1) for i in range(n)]
    length = {}
    for (l, r), nid in todo:
        if not (isinstance(nid, int) and 256 <= nid < 65535) or nid in length or not (0 <= l < 65535 and 0 <= r < 65535):
            raise ValueError("track_encoding: merged ids must be distinct ints in [256, 65535)")
        length[nid] = length.get(l, 1) + length.get(r, 1)
    tr = Trainer(n, len(todo))
    tr.load(data)
    tr.apply_pairs([p for p, _ in todo], [nid for _, nid in todo])
    ids = tr.ids()
    lens = np.ones(65536, np.int64)
    for nid, ln in length.items():
        lens[nid] = ln
    ends = np.cumsum(lens[ids])
    starts = ends - lens[ids]
    return ids.tolist(), list(zip(starts.tolist(), ends.tolist()))
