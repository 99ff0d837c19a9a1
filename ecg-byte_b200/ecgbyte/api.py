"""Tensor-in / tensor-out front end of libecgbyte.so.

torch is used for device memory, streams and (in dist_train.py) collectives only; all
tokenizer compute happens in the library's CUDA kernels.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib

_DT = {torch.float32: _lib.F32, torch.float64: _lib.F64, torch.int16: _lib.I16}
_NP_DT = {np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64, np.dtype(np.int16): _lib.I16}


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _np(a):
    return a.ctypes.data_as(C.c_void_p)


def _dev_index(device):
    if device is None:
        _lib.require_device()
        return torch.cuda.current_device()
    d = torch.device(device)
    if d.type != "cuda":
        raise ValueError("ecgbyte runs on CUDA devices only (got %s); there is no CPU path" % d)
    return d.index if d.index is not None else torch.cuda.current_device()


class Quantizer:
    """normalize_all's quantiser (tokenizer_utils.py:14-19) for one stored dtype."""

    def __init__(self, percentiles, dtype=torch.float32, device=None, i16_scale=1e-3):
        if isinstance(dtype, np.dtype) or dtype in (np.float32, np.float64, np.int16):
            dtype = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
                     np.dtype(np.int16): torch.int16}[np.dtype(dtype)]
        if dtype not in _DT:
            raise TypeError("unsupported sample dtype %r" % (dtype,))
        self.device = _dev_index(device)
        self.dtype = dtype
        self.p1 = float(percentiles["percentile_1"])
        self.p99 = float(percentiles["percentile_99"])
        self.i16_scale = float(i16_scale)
        h = C.c_void_p()
        check(lib().ecgb_quantizer_create(self.p1, self.p99, _DT[dtype], self.i16_scale, self.device, C.byref(h)))
        self._h = h

    def thresholds(self):
        out = np.zeros(25, np.float64)
        check(lib().ecgb_quantizer_thresholds(self._h, _np(out)))
        return out

    def _check_in(self, x):
        if not (isinstance(x, torch.Tensor) and x.is_cuda):
            raise TypeError("expected a CUDA tensor")
        if x.dtype != self.dtype:
            raise TypeError("quantizer was built for %s, got %s" % (self.dtype, x.dtype))
        if x.device.index != self.device:
            raise ValueError("tensor on cuda:%d, quantizer on cuda:%d" % (x.device.index, self.device))
        return x.contiguous()

    def quantize(self, x, out=None, direct=False):
        """CUDA tensor of samples -> uint8 tensor of symbols 'a'..'z' (same shape)."""
        x = self._check_in(x)
        if out is None:
            out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
        fn = lib().ecgb_quantize_direct if direct else lib().ecgb_quantize
        check(fn(self._h, _ptr(x), x.numel(), _ptr(out), _stream(x.device)))
        return out

    def quantize_host(self, x):
        """numpy array in host memory -> numpy uint8 symbols (H2D + kernel + D2H)."""
        x = np.ascontiguousarray(x)
        if _NP_DT.get(x.dtype) != _DT[self.dtype]:
            raise TypeError("quantizer was built for %s, got %s" % (self.dtype, x.dtype))
        out = np.empty(x.shape, np.uint8)
        check(lib().ecgb_quantize_host(self._h, _np(x), x.size, _np(out)))
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ecgb_quantizer_destroy(self._h)
                self._h = None
        except Exception:
            pass


def flatten_merges(merges):
    """list[tuple[list[int], int]] (the reference pickle form) -> flat numpy arrays."""
    M = len(merges)
    off = np.zeros(M + 1, np.uint64)
    ids = np.zeros(max(M, 1), np.uint32)
    lens = np.fromiter((len(s) for s, _ in merges), dtype=np.int64, count=M)
    off[1:] = np.cumsum(lens)
    for i, (_, t) in enumerate(merges):
        ids[i] = t
    seq = np.zeros(max(int(off[M]), 1), np.uint32)
    k = 0
    for s, _ in merges:
        n = len(s)
        seq[k: k + n] = s
        k += n
    return seq, off, ids


class Vocab:
    """A merges table flattened into a device-resident trie (lib.rs:127-161)."""

    def __init__(self, merges=None, flat=None, device=None):
        self.device = _dev_index(device)
        seq, off, ids = flat if flat is not None else flatten_merges(merges)
        seq = np.ascontiguousarray(seq, np.uint32)
        off = np.ascontiguousarray(off, np.uint64)
        ids = np.ascontiguousarray(ids, np.uint32)
        self.n_merges = len(off) - 1
        h = C.c_void_p()
        check(lib().ecgb_vocab_create(_np(seq), _np(off), _np(ids), self.n_merges, self.device, C.byref(h)))
        self._h = h
        self.flat = (seq, off, ids)

    @classmethod
    def from_pairs(cls, pairs, device=None):
        seq, off = expand_merges(pairs)
        ids = np.arange(256, 256 + len(off) - 1, dtype=np.uint32)
        return cls(flat=(seq, off, ids), device=device)

    def info(self):
        inf = _lib.VocabInfo()
        check(lib().ecgb_vocab_info(self._h, C.byref(inf)))
        return {n: getattr(inf, n) for n, _ in inf._fields_}

    def encode_symbols(self, sym, out_stride=None, offsets=None, tokens=None, lens=None):
        """uint8 CUDA tensor [n_rec, rec_len] of text bytes (or flat + offsets) ->
        (tokens int32 [n_rec, out_stride], lens int32 [n_rec])."""
        if not (isinstance(sym, torch.Tensor) and sym.is_cuda and sym.dtype == torch.uint8):
            raise TypeError("expected a uint8 CUDA tensor")
        sym = sym.contiguous()
        if offsets is not None:
            offsets = offsets.to(device=sym.device, dtype=torch.int64).contiguous()
            n_rec, rec_len = offsets.numel() - 1, 0
            max_len = int((offsets[1:] - offsets[:-1]).max().item()) if n_rec else 0
        else:
            if sym.dim() == 1:
                sym = sym.unsqueeze(0)
            n_rec, rec_len = sym.shape[0], sym[0].numel() if sym.shape[0] else 0
            max_len = rec_len
        if out_stride is None:
            out_stride = max(max_len, 1)
        if tokens is None:
            tokens = torch.empty((n_rec, out_stride), dtype=torch.int32, device=sym.device)
        if lens is None:
            lens = torch.empty((n_rec,), dtype=torch.int32, device=sym.device)
        check(lib().ecgb_encode_symbols(self._h, _ptr(sym), n_rec, rec_len,
                                        _ptr(offsets) if offsets is not None else None,
                                        _ptr(tokens), out_stride, _ptr(lens), _stream(sym.device)))
        return tokens, lens

    def encode_batch(self, quantizer, x, out_stride=None, tokens=None, lens=None):
        """Fused quantise + encode: CUDA tensor [n_rec, C, L] (or [n_rec, C*L]) of raw
        samples -> (tokens int32 [n_rec, out_stride], lens int32 [n_rec])."""
        x = quantizer._check_in(x)
        n_rec = x.shape[0]
        rec_len = x[0].numel() if n_rec else 0
        if out_stride is None:
            out_stride = max(rec_len, 1)
        if tokens is None:
            tokens = torch.empty((n_rec, out_stride), dtype=torch.int32, device=x.device)
        if lens is None:
            lens = torch.empty((n_rec,), dtype=torch.int32, device=x.device)
        check(lib().ecgb_encode_batch(self._h, quantizer._h, _ptr(x), n_rec, rec_len, _ptr(tokens), out_stride,
                                      _ptr(lens), _stream(x.device)))
        return tokens, lens

    def encode_batch_host(self, quantizer, x, out_stride):
        """numpy samples in host memory -> numpy tokens / lens, copies included."""
        x = np.ascontiguousarray(x)
        n_rec = x.shape[0]
        rec_len = x[0].size if n_rec else 0
        tokens = np.empty((n_rec, out_stride), np.int32)
        lens = np.empty((n_rec,), np.int32)
        check(lib().ecgb_encode_batch_host(self._h, quantizer._h, _np(x), n_rec, rec_len, _np(tokens), out_stride,
                                           _np(lens)))
        return tokens, lens

    def decode_symbols(self, tokens, lens, sym_stride):
        """decode_text (tu.py:75-77) for a batch: int32 CUDA tokens [n, stride] + lens [n] ->
        (uint8 symbols [n, sym_stride], int32 sym_len [n])."""
        tokens = tokens.contiguous()
        lens = lens.contiguous()
        n = tokens.shape[0]
        sym = torch.empty((n, sym_stride), dtype=torch.uint8, device=tokens.device)
        sym_len = torch.empty((n,), dtype=torch.int32, device=tokens.device)
        check(lib().ecgb_decode_symbols(self._h, _ptr(tokens), n, tokens.shape[1], _ptr(lens), _ptr(sym), sym_stride,
                                        _ptr(sym_len), _stream(tokens.device)))
        return sym, sym_len

    def expand_attention(self, tokens, lens, attn, out_stride):
        """expand_attention (runners/interpret.py:106-111) for a batch: float32 attention per token
        [n, stride] -> per base symbol [n, out_stride] (+ int32 lengths [n])."""
        tokens = tokens.contiguous()
        lens = lens.contiguous()
        attn = attn.to(torch.float32).contiguous()
        assert attn.shape == tokens.shape
        n = tokens.shape[0]
        out = torch.zeros((n, out_stride), dtype=torch.float32, device=tokens.device)
        out_len = torch.empty((n,), dtype=torch.int32, device=tokens.device)
        check(lib().ecgb_expand_attention(self._h, _ptr(tokens), _ptr(attn), n, tokens.shape[1], _ptr(lens), _ptr(out),
                                          out_stride, _ptr(out_len), _stream(tokens.device)))
        return out, out_len

    def encode_text(self, text):
        """One string / bytes object -> list[int] (rust_bpe.encode_text semantics)."""
        data = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        n = len(data)
        if n == 0:
            return []
        buf = np.frombuffer(data, np.uint8)
        out = np.empty(n, np.uint32)
        n_out = C.c_size_t(0)
        check(lib().ecgb_encode_text_host(self._h, _np(buf), n, _np(out), n, C.byref(n_out)))
        return out[: n_out.value].tolist()

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ecgb_vocab_destroy(self._h)
                self._h = None
        except Exception:
            pass


def token_histogram(tokens, lens, n_ids, counts=None):
    """Counter over the encoded ids of a batch (analyze_token_distribution, tu.py:44-49): int32 CUDA
    tokens [n, stride] + lens [n] -> int64 CUDA counts [n_ids] (accumulated into `counts` if given)."""
    tokens = tokens.contiguous()
    lens = lens.contiguous()
    if counts is None:
        counts = torch.zeros((n_ids,), dtype=torch.int64, device=tokens.device)
    assert counts.dtype == torch.int64 and counts.numel() == n_ids and counts.is_contiguous()
    dev = tokens.device.index if tokens.device.index is not None else torch.cuda.current_device()
    check(lib().ecgb_token_histogram(_ptr(tokens), tokens.shape[1], _ptr(lens), tokens.shape[0], n_ids, _ptr(counts), dev,
                                     _stream(tokens.device)))
    return counts


def minmax(x):
    """np.min / np.max of a CUDA tensor (float32 / float64 / int16), NaN propagates."""
    x = x.contiguous()
    lo, hi = C.c_double(0), C.c_double(0)
    check(lib().ecgb_minmax(_ptr(x), _DT[x.dtype], x.numel(), C.byref(lo), C.byref(hi), x.device.index, _stream(x.device)))
    return lo.value, hi.value


def percentiles(samples, q):
    """np.percentile(samples, q) (method 'linear') for a float64 CUDA tensor of samples."""
    s = samples.to(torch.float64).contiguous()
    qs = np.ascontiguousarray(np.atleast_1d(q), np.float64)
    out = np.zeros(qs.size, np.float64)
    check(lib().ecgb_percentiles(_ptr(s), s.numel(), _np(qs), qs.size, _np(out), s.device.index, _stream(s.device)))
    return out


def global_stats(segments, samples, skipped_instances=0):
    """The stats dict of compute_global_stats (preprocess_utils.py:168-213): min / max over all stored
    samples of `segments`, 1st / 99th percentile of `samples` (the reference draws ~100k of them)."""
    lo, hi = minmax(segments)
    p = percentiles(samples, [1, 99])
    return {"global_min": np.float64(lo), "global_max": np.float64(hi), "percentile_1": np.float64(p[0]),
            "percentile_99": np.float64(p[1]), "skipped_instances": skipped_instances}


def dequantize(sym, percentiles):
    """reverse_normalize_all (tu.py:22-28): uint8 CUDA symbols -> float64 CUDA values."""
    sym = sym.contiguous()
    out = torch.empty(sym.shape, dtype=torch.float64, device=sym.device)
    check(lib().ecgb_dequantize(float(percentiles["percentile_1"]), float(percentiles["percentile_99"]), _ptr(sym),
                                sym.numel(), _ptr(out), sym.device.index, _stream(sym.device)))
    return out


def pack_training(tokens, lens, lut, text, text_off, q_len, pad_to_max, pad_id, bos_id, eos_id, sig_start_id, sig_end_id):
    """ECGTokenDataset post-processing (data_loader.py:80, 101-132) for a batch on the device.
    tokens int32 [n, stride], lens int32 [n] (encoder output); lut int64 [256 + M] = LLM id of
    'signal_k'; text int64 flat question+answer ids, text_off int64 [n + 1], q_len int32 [n].
    Returns (input_ids, attn_mask, labels, position_ids, status), rows of pad_to_max + 4."""
    dev = tokens.device
    n = tokens.shape[0]
    P = int(pad_to_max) + 4
    ids = torch.empty((n, P), dtype=torch.int64, device=dev)
    attn = torch.empty((n, P), dtype=torch.float32, device=dev)
    labels = torch.empty((n, P), dtype=torch.int64, device=dev)
    pos = torch.empty((n, P), dtype=torch.int64, device=dev)
    status = torch.empty((n,), dtype=torch.int32, device=dev)
    cfg = _lib.PackCfg(int(pad_id), int(bos_id), int(eos_id), int(sig_start_id), int(sig_end_id), int(pad_to_max))
    tokens, lens = tokens.contiguous(), lens.contiguous()
    lut = lut.to(device=dev, dtype=torch.int64).contiguous()
    text = text.to(device=dev, dtype=torch.int64).contiguous()
    text_off = text_off.to(device=dev, dtype=torch.int64).contiguous()
    q_len = q_len.to(device=dev, dtype=torch.int32).contiguous()
    check(lib().ecgb_pack_training(_ptr(tokens), tokens.shape[1], _ptr(lens), n, _ptr(lut), lut.numel(), _ptr(text),
                                   _ptr(text_off), _ptr(q_len), C.byref(cfg), _ptr(ids), _ptr(attn), _ptr(labels),
                                   _ptr(pos), _ptr(status), dev.index, _stream(dev)))
    return ids, attn, labels, pos, status


class EncodePipeline:
    """Host-buffer front end of the fused encoder: records in PINNED host memory ->
    tokens / lengths in pinned host memory.  Chunks are double-buffered over CUDA streams
    so the H2D copy of chunk i+1, the kernel of chunk i and the D2H copy of chunk i-1
    overlap (the path is PCIe-bound: 240 KB in per fp32 record)."""

    def __init__(self, vocab, quantizer, rec_len, out_stride, chunk=4096, depth=4):
        self.vocab, self.q = vocab, quantizer
        self.rec_len, self.out_stride, self.chunk, self.depth = rec_len, out_stride, chunk, depth
        dev = torch.device("cuda", vocab.device)
        self.dev = dev
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
        self.d_in = [torch.empty((chunk, rec_len), dtype=quantizer.dtype, device=dev) for _ in range(depth)]
        self.d_tok = [torch.empty((chunk, out_stride), dtype=torch.int32, device=dev) for _ in range(depth)]
        self.d_len = [torch.empty((chunk,), dtype=torch.int32, device=dev) for _ in range(depth)]

    def run(self, x_pinned, tokens_pinned, lens_pinned, after_current_stream=False):
        """Enqueues the whole batch and returns; the caller's current stream waits for the results (synchronise it, or
        record an event on it, before reading the pinned outputs).  Consecutive calls pipeline into each other: the
        first chunks of the next batch are copied in while the last chunks of this one are still being encoded and
        copied out (a chunk's kernel has a latency floor of several ms -- one walker per record -- which would
        otherwise be paid as a drain at the end of every call).  after_current_stream=True orders the batch behind
        work already enqueued on the current stream (inputs produced on the device side)."""
        n = x_pinned.shape[0]
        x2 = x_pinned.view(n, self.rec_len)
        cur = torch.cuda.current_stream(self.dev)
        if after_current_stream:
            for s in self.streams:
                s.wait_stream(cur)
        k = 0
        for c0 in range(0, n, self.chunk):
            m = min(self.chunk, n - c0)
            b = k % self.depth
            k += 1
            with torch.cuda.stream(self.streams[b]):
                self.d_in[b][:m].copy_(x2[c0:c0 + m], non_blocking=True)
                self.vocab.encode_batch(self.q, self.d_in[b][:m], out_stride=self.out_stride,
                                        tokens=self.d_tok[b], lens=self.d_len[b])
                tokens_pinned[c0:c0 + m].copy_(self.d_tok[b][:m], non_blocking=True)
                lens_pinned[c0:c0 + m].copy_(self.d_len[b][:m], non_blocking=True)
        for s in self.streams:
            cur.wait_stream(s)
        return k  # kernel launches


class EncodePipelineCSR(EncodePipeline):
    """EncodePipeline with compact output: what crosses PCIe on the way back is 2 bytes per token plus 12 bytes per
    record (offset, length) instead of a padded int32 row of out_stride slots.  Record r's tokens are
    tokens16[off[r] : off[r] + lens[r]]; rows are back to back inside a chunk of `chunk` records, and chunk k starts at
    k * chunk * out_stride (no size has to travel to the host before the data does).

    The compaction kernel stores the tokens STRAIGHT into the pinned host buffer (pinned memory is mapped into the
    device's address space), so there is no device-to-host token copy to size and no host synchronisation anywhere in
    the pipeline."""

    def __init__(self, vocab, quantizer, rec_len, out_stride, chunk=4096, depth=4):
        super().__init__(vocab, quantizer, rec_len, out_stride, chunk, depth)
        self.d_off = [torch.empty((chunk + 1,), dtype=torch.int64, device=self.dev) for _ in range(depth)]

    def run(self, x_pinned, tokens16_pinned, lens_pinned, off_pinned, after_current_stream=False):
        """tokens16_pinned: flat pinned uint16 buffer of n * out_stride slots; lens_pinned int32 [n]; off_pinned int64 [n].
        Stream semantics as EncodePipeline.run (consecutive calls pipeline into each other).  Returns the number of
        kernel launches."""
        n = x_pinned.shape[0]
        if tokens16_pinned.numel() < n * self.out_stride:
            raise ValueError("tokens16 buffer needs n * out_stride slots")
        if not (tokens16_pinned.is_pinned() and lens_pinned.is_pinned() and off_pinned.is_pinned()):
            raise ValueError("outputs must be pinned host tensors")
        x2 = x_pinned.view(n, self.rec_len)
        cur = torch.cuda.current_stream(self.dev)
        if after_current_stream:
            for s in self.streams:
                s.wait_stream(cur)
        k = 0
        for c0 in range(0, n, self.chunk):
            m = min(self.chunk, n - c0)
            b = k % self.depth
            k += 1
            with torch.cuda.stream(self.streams[b]):
                self.d_in[b][:m].copy_(x2[c0:c0 + m], non_blocking=True)
                self.vocab.encode_batch(self.q, self.d_in[b][:m], out_stride=self.out_stride,
                                        tokens=self.d_tok[b], lens=self.d_len[b])
                check(lib().ecgb_tokens_csr(_ptr(self.d_tok[b]), self.out_stride, _ptr(self.d_len[b]), m,
                                            _ptr(tokens16_pinned), _ptr(self.d_off[b]), c0 * self.out_stride,
                                            self.dev.index, _stream(self.dev)))
                off_pinned[c0:c0 + m].copy_(self.d_off[b][:m], non_blocking=True)
                lens_pinned[c0:c0 + m].copy_(self.d_len[b][:m], non_blocking=True)
        for s in self.streams:
            cur.wait_stream(s)
        return 3 * k


def expand_merges(pairs):
    """pairs [M, 2] -> (seq u32, off u64[M+1]): the expanded sequences (lib.rs:101-110)."""
    pairs = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
    M = pairs.shape[0]
    off = np.zeros(M + 1, np.uint64)
    rc = lib().ecgb_expand_merges(_np(pairs), M, None, 0, _np(off))
    if rc not in (_lib.OK, _lib.ECAPACITY):
        check(rc)
    seq = np.zeros(max(int(off[M]), 1), np.uint32)
    check(lib().ecgb_expand_merges(_np(pairs), M, _np(seq), seq.size, _np(off)))
    return seq[: int(off[M])], off


class Trainer:
    """byte_pair_encoding (lib.rs:58-125) on one device (or one shard of a corpus)."""

    STEP_DEVICE = 0xFFFFFFFF  # dist_commit / dist_merge: step number from the device-side counter

    def __init__(self, capacity_tokens, max_merges, device=None, table_log2=0):
        self.device = _dev_index(device)
        self.max_merges = int(max_merges)
        self.capacity = int(capacity_tokens)
        h = C.c_void_p()
        check(lib().ecgb_trainer_create(self.device, self.capacity, self.max_merges, int(table_log2), C.byref(h)))
        self._h = h

    def load(self, text):
        """text: bytes / numpy uint8 (host) or uint8 CUDA tensor."""
        if isinstance(text, torch.Tensor):
            if not (text.is_cuda and text.dtype == torch.uint8):
                raise TypeError("expected a uint8 CUDA tensor")
            text = text.contiguous()
            check(lib().ecgb_trainer_load_device(self._h, _ptr(text), text.numel(), _stream(text.device)))
            torch.cuda.current_stream(text.device).synchronize()
            return
        if isinstance(text, str):
            text = text.encode("utf-8")
        buf = np.frombuffer(bytes(text), np.uint8) if isinstance(text, (bytes, bytearray)) else \
            np.ascontiguousarray(text, np.uint8)
        check(lib().ecgb_trainer_load_host(self._h, _np(buf), buf.size))

    def run(self, num_merges):
        """-> (pairs u32 [m, 2], counts u64 [m], ntied u32 [m])"""
        m = int(num_merges)
        pairs = np.zeros((max(m, 1), 2), np.uint32)
        counts = np.zeros(max(m, 1), np.uint64)
        ntied = np.zeros(max(m, 1), np.uint32)
        done = C.c_uint32(0)
        check(lib().ecgb_trainer_run(self._h, m, _np(pairs), _np(counts), _np(ntied), C.byref(done)))
        d = done.value
        return pairs[:d].copy(), counts[:d].copy(), ntied[:d].copy()

    def length(self):
        n = C.c_uint64(0)
        check(lib().ecgb_trainer_length(self._h, C.byref(n)))
        return n.value

    def ids(self):
        n = self.length()
        out = np.empty(max(n, 1), np.uint32)
        got = C.c_uint64(0)
        check(lib().ecgb_trainer_ids_host(self._h, _np(out), out.size, C.byref(got)))
        return out[: got.value]

    def lengths(self, n_steps):
        """stream length after 0..n_steps merges (uint64 [n_steps + 1])."""
        out = np.zeros(int(n_steps) + 1, np.uint64)
        check(lib().ecgb_trainer_lengths(self._h, int(n_steps), _np(out)))
        return out

    def apply_pairs(self, pairs, new_ids):
        """Apply the merges (left, right) -> new id, in order, to the loaded text (no training)."""
        pairs = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
        new_ids = np.ascontiguousarray(new_ids, np.uint32).reshape(-1)
        assert len(pairs) == len(new_ids)
        check(lib().ecgb_trainer_apply_pairs(self._h, _np(pairs), _np(new_ids), len(new_ids)))

    def table_stats(self):
        """{'used', 'capacity', 'candidates', 'overflow'} of the pair table."""
        out = np.zeros(4, np.uint64)
        check(lib().ecgb_trainer_table_stats(self._h, _np(out)))
        return {"used": int(out[0]), "capacity": int(out[1]), "candidates": int(out[2]), "overflow": int(out[3])}

    def histogram(self):
        """{(left, right): count} of the live pair histogram."""
        n = C.c_uint64(0)
        rc = lib().ecgb_trainer_histogram(self._h, None, None, 0, C.byref(n))
        if rc not in (_lib.OK, _lib.ECAPACITY):
            check(rc)
        pairs = np.zeros((max(n.value, 1), 2), np.uint32)
        counts = np.zeros(max(n.value, 1), np.int64)
        check(lib().ecgb_trainer_histogram(self._h, _np(pairs), _np(counts), pairs.shape[0], C.byref(n)))
        return {(int(l), int(r)): int(c) for (l, r), c in zip(pairs[: n.value].tolist(), counts[: n.value].tolist())}

    # ---- sharded interface (see dist_train.py) ----
    def dist_sizes(self):
        b, l = C.c_uint32(0), C.c_uint32(0)
        check(lib().ecgb_trainer_dist_sizes(self._h, C.byref(b), C.byref(l)))
        return b.value, l.value

    def dist_begin(self, rank, world, boundary_out):
        check(lib().ecgb_trainer_dist_begin(self._h, rank, world, _ptr(boundary_out), _stream(boundary_out.device)))

    def dist_count(self, all_boundaries, list_out):
        check(lib().ecgb_trainer_dist_count(self._h, _ptr(all_boundaries), _ptr(list_out), _stream(list_out.device)))

    def dist_commit(self, step, all_lists, boundary_out):
        check(lib().ecgb_trainer_dist_commit(self._h, step, _ptr(all_lists), _ptr(boundary_out),
                                             _stream(boundary_out.device)))

    def dist_merge(self, step, all_boundaries, list_out):
        check(lib().ecgb_trainer_dist_merge(self._h, step, _ptr(all_boundaries), _ptr(list_out),
                                            _stream(list_out.device)))

    def dist_advance(self, device):
        """Increment the device-side step counter (see STEP_DEVICE)."""
        check(lib().ecgb_trainer_dist_advance(self._h, _stream(device)))

    # ---- persistent sharded loop: device-initiated exchange over peer memory ----
    def peer_area(self, world):
        """Allocates / clears this rank's receive area -> (device pointer, bytes)."""
        p, n = C.c_void_p(), C.c_uint64(0)
        check(lib().ecgb_trainer_peer_area(self._h, int(world), C.byref(p), C.byref(n)))
        return p.value, n.value

    def dist_apply(self, all_lists, world):
        check(lib().ecgb_trainer_dist_apply(self._h, _ptr(all_lists), int(world), _stream(all_lists.device)))

    def dist_run(self, rank, world, areas, all_boundaries, num_merges, max_ctas=0, timeout_s=0.0):
        """areas: device pointers (ints) of every rank's receive area as addressable here."""
        arr = (C.c_void_p * int(world))(*[C.c_void_p(a) for a in areas])
        check(lib().ecgb_trainer_dist_run(self._h, int(rank), int(world), arr, _ptr(all_boundaries), int(num_merges),
                                          int(max_ctas), float(timeout_s), _stream(all_boundaries.device)))

    def results(self, n_steps):
        m = int(n_steps)
        pairs = np.zeros((max(m, 1), 2), np.uint32)
        counts = np.zeros(max(m, 1), np.uint64)
        ntied = np.zeros(max(m, 1), np.uint32)
        done = C.c_uint32(0)
        check(lib().ecgb_trainer_results(self._h, m, _np(pairs), _np(counts), _np(ntied), C.byref(done)))
        d = done.value
        return pairs[:d].copy(), counts[:d].copy(), ntied[:d].copy()

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().ecgb_trainer_destroy(self._h)
                self._h = None
        except Exception:
            pass
