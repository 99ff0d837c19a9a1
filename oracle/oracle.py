"""ctypes front-end of the CPU oracle (oracle/ecgb_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing under ecg-byte_b200/
imports this module.

Return types mirror the reference's Python-visible types
(/root/reference/ecg_byte/rust_bpe/src/lib.rs:58-63,149-150).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libecgb_oracle.so")
_lib = None

DT_F32, DT_F64, DT_I16 = 0, 1, 2
_DT = {np.dtype(np.float32): DT_F32, np.dtype(np.float64): DT_F64, np.dtype(np.int16): DT_I16}


def build(force=False):
    src = os.path.join(_HERE, "ecgb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libecgb_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, sz, u32, u64, dbl = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_double
        L.ecgo_quantize.argtypes = [vp, C.c_int, sz, dbl, dbl, dbl, vp]
        L.ecgo_train_naive.argtypes = [vp, sz, u32, C.c_int, vp, vp, vp, vp, vp, vp]
        L.ecgo_train_fast.argtypes = [vp, sz, u32, vp, vp, vp, vp, vp, vp]
        L.ecgo_expand_merges.argtypes = [vp, u32, vp, u64, vp]
        L.ecgo_trie_build.argtypes = [vp, vp, vp, u32]
        L.ecgo_trie_build.restype = vp
        L.ecgo_trie_free.argtypes = [vp]
        L.ecgo_trie_free.restype = None
        L.ecgo_trie_nodes.argtypes = [vp]
        L.ecgo_trie_nodes.restype = u64
        L.ecgo_trie_encode.argtypes = [vp, vp, sz, vp, sz, vp]
        L.ecgo_encode.argtypes = [vp, sz, vp, vp, vp, u32, vp, sz, vp]
        L.ecgo_trie_encode_batch.argtypes = [vp, vp, sz, sz, vp, sz, vp]
        L.ecgo_decode.argtypes = [vp, sz, vp, vp, u32, vp, sz, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------- quantise
def quantize(signal, p1, p99, i16_scale=1e-3):
    """normalize_all's symbol output as uint8 codes 'a'..'z' (tu.py:14-19)."""
    x = np.ascontiguousarray(signal)
    out = np.empty(x.shape, np.uint8)
    rc = lib().ecgo_quantize(_p(x), _DT[x.dtype], x.size, float(p1), float(p99), float(i16_scale), _p(out))
    assert rc == 0, rc
    return out


# ------------------------------------------------------------------------ train
def _as_bytes(text):
    if isinstance(text, str):
        text = text.encode("utf-8")
    if isinstance(text, (bytes, bytearray)):
        return np.frombuffer(bytes(text), np.uint8)
    return np.ascontiguousarray(text, np.uint8)


def train_pairs(text, num_merges, num_threads=1, fast=False):
    """-> (ids u32[n'], pairs u32[M,2], counts u64[M], ntied u32[M])"""
    t = _as_bytes(text)
    n = t.size
    ids = np.empty(max(n, 1), np.uint32)
    pairs = np.zeros((max(num_merges, 1), 2), np.uint32)
    counts = np.zeros(max(num_merges, 1), np.uint64)
    ntied = np.zeros(max(num_merges, 1), np.uint32)
    n_ids = C.c_size_t(0)
    done = C.c_uint32(0)
    if fast:
        rc = lib().ecgo_train_fast(_p(t), n, num_merges, _p(ids), C.byref(n_ids), _p(pairs), _p(counts),
                                   _p(ntied), C.byref(done))
    else:
        rc = lib().ecgo_train_naive(_p(t), n, num_merges, num_threads, _p(ids), C.byref(n_ids), _p(pairs),
                                    _p(counts), _p(ntied), C.byref(done))
    assert rc == 0, rc
    m = done.value
    return ids[: n_ids.value].copy(), pairs[:m].copy(), counts[:m].copy(), ntied[:m].copy()


def expand(pairs):
    """pairs u32[M,2] -> (seq u32[total], off u64[M+1]) (lib.rs:101-110)."""
    pairs = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
    M = pairs.shape[0]
    off = np.zeros(M + 1, np.uint64)
    seq = np.zeros(1, np.uint32)
    rc = lib().ecgo_expand_merges(_p(pairs), M, _p(seq), 0, _p(off))
    if rc == 3:
        seq = np.zeros(int(off[M]), np.uint32)
        rc = lib().ecgo_expand_merges(_p(pairs), M, _p(seq), seq.size, _p(off))
    assert rc == 0, rc
    return seq[: int(off[M])], off


def byte_to_string(b):
    """lib.rs:50-56"""
    return chr(b) if b <= 127 else "<%d>" % b


def to_reference_types(ids, pairs):
    """(ids, pairs) -> the tuple rust_bpe.byte_pair_encoding returns (lib.rs:124)."""
    seq, off = expand(pairs)
    vocab = {i: byte_to_string(i) for i in range(256)}
    merges = []
    for i, (l, r) in enumerate(np.asarray(pairs).tolist()):
        vocab[256 + i] = vocab[l] + vocab[r]
        merges.append((seq[int(off[i]): int(off[i + 1])].tolist(), 256 + i))
    return [int(v) for v in ids], vocab, merges


def byte_pair_encoding(text, num_merges, num_threads=1, fast=False):
    ids, pairs, _, _ = train_pairs(text, num_merges, num_threads, fast)
    return to_reference_types(ids, pairs)


# ----------------------------------------------------------------------- encode
def flatten_merges(merges):
    """list[(list[int], int)] -> (seq u32, off u64[M+1], ids u32[M])"""
    M = len(merges)
    off = np.zeros(M + 1, np.uint64)
    ids = np.zeros(max(M, 1), np.uint32)
    tot = 0
    for i, (s, t) in enumerate(merges):
        tot += len(s)
        off[i + 1] = tot
        ids[i] = t
    seq = np.zeros(max(tot, 1), np.uint32)
    k = 0
    for s, _ in merges:
        seq[k: k + len(s)] = s
        k += len(s)
    return seq, off, ids[:M] if M else ids[:0]


def encode_text(text, merges):
    """rust_bpe.encode_text semantics: trie rebuilt per call (lib.rs:149-193)."""
    t = _as_bytes(text)
    seq, off, ids = flatten_merges(merges)
    idsb = np.ascontiguousarray(ids, np.uint32) if len(ids) else np.zeros(1, np.uint32)
    out = np.empty(max(t.size, 1), np.uint32)
    n_out = C.c_size_t(0)
    rc = lib().ecgo_encode(_p(t), t.size, _p(seq), _p(off), _p(idsb), len(merges), _p(out), out.size,
                           C.byref(n_out))
    assert rc == 0, rc
    return out[: n_out.value].tolist()


class Trie:
    """Trie built once (amortised baseline mode)."""

    def __init__(self, merges=None, flat=None):
        seq, off, ids = flat if flat is not None else flatten_merges(merges)
        self._keep = (seq, off, np.ascontiguousarray(ids, np.uint32) if len(ids) else np.zeros(1, np.uint32))
        self.M = len(off) - 1
        self.h = lib().ecgo_trie_build(_p(self._keep[0]), _p(self._keep[1]), _p(self._keep[2]), self.M)
        assert self.h

    @property
    def nodes(self):
        return int(lib().ecgo_trie_nodes(self.h))

    def encode(self, sym):
        t = _as_bytes(sym)
        out = np.empty(max(t.size, 1), np.uint32)
        n_out = C.c_size_t(0)
        rc = lib().ecgo_trie_encode(self.h, _p(t), t.size, _p(out), out.size, C.byref(n_out))
        assert rc == 0, rc
        return out[: n_out.value].copy()

    def encode_batch(self, sym2d, stride=None):
        s = np.ascontiguousarray(sym2d, np.uint8)
        n_rec, rec_len = s.shape
        stride = stride or rec_len
        out = np.zeros((n_rec, stride), np.uint32)
        lens = np.zeros(n_rec, np.uint32)
        lib().ecgo_trie_encode_batch(self.h, _p(s), n_rec, rec_len, _p(out), stride, _p(lens))
        return out, lens

    def __del__(self):
        try:
            lib().ecgo_trie_free(self.h)
        except Exception:
            pass


def decode(tokens, merges=None, flat=None):
    seq, off, _ = flat if flat is not None else flatten_merges(merges)
    tk = np.ascontiguousarray(tokens, np.uint32)
    n_out = C.c_size_t(0)
    out = np.empty(1, np.uint8)
    rc = lib().ecgo_decode(_p(tk), tk.size, _p(seq), _p(off), len(off) - 1, _p(out), 0, C.byref(n_out))
    assert rc in (0, 3), rc
    out = np.empty(max(n_out.value, 1), np.uint8)
    rc = lib().ecgo_decode(_p(tk), tk.size, _p(seq), _p(off), len(off) - 1, _p(out), out.size, C.byref(n_out))
    assert rc == 0, rc
    return out[: n_out.value]
