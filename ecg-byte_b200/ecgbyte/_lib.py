"""ctypes binding of libecgbyte.so (include/ecgbyte.h).

The library is the product: if it cannot be loaded, or no CUDA device is usable,
every compute call raises -- there is no CPU fallback in this package.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_PKG, "lib", "libecgbyte.so")

OK, EINVAL, ENOMEM, ECUDA, ENODEVICE, ECAPACITY, EUNSUPPORTED = range(7)
F32, F64, I16, U8 = range(4)


class EcgbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("libecgbyte status %d: %s" % (status, msg))
        self.status = status


class PackCfg(C.Structure):
    _fields_ = [("pad_id", C.c_int64), ("bos_id", C.c_int64), ("eos_id", C.c_int64), ("sig_start_id", C.c_int64),
                ("sig_end_id", C.c_int64), ("pad_to_max", C.c_uint32)]


class VocabInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "n_merges", "n_nodes", "n_classes", "compact", "max_token_len", "node_bytes", "smem_nodes", "pair_slots")]


_lib = None


def _declare(L):
    vp, sz, u32, u64, i32, dbl = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_int, C.c_double
    pp = C.POINTER(C.c_void_p)
    sig = {
        "ecgb_last_error": ([], C.c_char_p),
        "ecgb_version": ([], i32),
        "ecgb_device_count": ([C.POINTER(i32)], i32),
        "ecgb_quantizer_create": ([dbl, dbl, i32, dbl, i32, pp], i32),
        "ecgb_quantizer_destroy": ([vp], i32),
        "ecgb_quantizer_thresholds": ([vp, vp], i32),
        "ecgb_quantize": ([vp, vp, sz, vp, vp], i32),
        "ecgb_quantize_direct": ([vp, vp, sz, vp, vp], i32),
        "ecgb_quantize_host": ([vp, vp, sz, vp], i32),
        "ecgb_vocab_create": ([vp, vp, vp, u32, i32, pp], i32),
        "ecgb_vocab_destroy": ([vp], i32),
        "ecgb_vocab_info": ([vp, C.POINTER(VocabInfo)], i32),
        "ecgb_pairtab_host": ([vp, vp, vp, u32, vp, vp, u32, C.POINTER(u32), vp, vp], i32),
        "ecgb_encode_symbols": ([vp, vp, sz, sz, vp, vp, sz, vp, vp], i32),
        "ecgb_encode_batch": ([vp, vp, vp, sz, sz, vp, sz, vp, vp], i32),
        "ecgb_encode_text_host": ([vp, vp, sz, vp, sz, C.POINTER(sz)], i32),
        "ecgb_encode_batch_host": ([vp, vp, vp, sz, sz, vp, sz, vp], i32),
        "ecgb_decode_symbols": ([vp, vp, sz, sz, vp, vp, sz, vp, vp], i32),
        "ecgb_dequantize": ([dbl, dbl, vp, sz, vp, i32, vp], i32),
        "ecgb_expand_attention": ([vp, vp, vp, sz, sz, vp, vp, sz, vp, vp], i32),
        "ecgb_tokens_csr": ([vp, sz, vp, sz, vp, vp, u64, i32, vp], i32),
        "ecgb_token_histogram": ([vp, sz, vp, sz, u32, vp, i32, vp], i32),
        "ecgb_minmax": ([vp, i32, sz, C.POINTER(dbl), C.POINTER(dbl), i32, vp], i32),
        "ecgb_percentiles": ([vp, sz, vp, i32, vp, i32, vp], i32),
        "ecgb_pack_training": ([vp, sz, vp, sz, vp, u32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp], i32),
        "ecgb_trainer_create": ([i32, u64, u32, u32, pp], i32),
        "ecgb_trainer_destroy": ([vp], i32),
        "ecgb_trainer_load_device": ([vp, vp, u64, vp], i32),
        "ecgb_trainer_load_host": ([vp, vp, u64], i32),
        "ecgb_trainer_run": ([vp, u32, vp, vp, vp, C.POINTER(u32)], i32),
        "ecgb_trainer_length": ([vp, C.POINTER(u64)], i32),
        "ecgb_trainer_ids_host": ([vp, vp, u64, C.POINTER(u64)], i32),
        "ecgb_trainer_lengths": ([vp, u32, vp], i32),
        "ecgb_trainer_apply_pairs": ([vp, vp, vp, u32], i32),
        "ecgb_trainer_table_stats": ([vp, vp], i32),
        "ecgb_trainer_histogram": ([vp, vp, vp, u64, C.POINTER(u64)], i32),
        "ecgb_trainer_dist_sizes": ([vp, C.POINTER(u32), C.POINTER(u32)], i32),
        "ecgb_trainer_dist_begin": ([vp, i32, i32, vp, vp], i32),
        "ecgb_trainer_dist_count": ([vp, vp, vp, vp], i32),
        "ecgb_trainer_dist_commit": ([vp, u32, vp, vp, vp], i32),
        "ecgb_trainer_dist_merge": ([vp, u32, vp, vp, vp], i32),
        "ecgb_trainer_dist_advance": ([vp, vp], i32),
        "ecgb_trainer_results": ([vp, u32, vp, vp, vp, C.POINTER(u32)], i32),
        "ecgb_trainer_peer_area": ([vp, i32, pp, C.POINTER(u64)], i32),
        "ecgb_trainer_dist_apply": ([vp, vp, i32, vp], i32),
        "ecgb_trainer_dist_run": ([vp, i32, i32, pp, vp, u32, u32, dbl, vp], i32),
        "ecgb_trainer_dist_run_local": ([pp, i32, pp, vp, u32, u32, dbl, vp], i32),
        "ecgb_ipc_export": ([vp, vp], i32),
        "ecgb_ipc_open": ([vp, i32, pp], i32),
        "ecgb_ipc_close": ([vp, i32], i32),
        "ecgb_expand_merges": ([vp, u32, vp, u64, vp], i32),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    return sig


EXPORTS = None


def lib():
    """Loads the CUDA library (fails loudly if it is missing)."""
    global _lib, EXPORTS
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libecgbyte.so is not built (%s). Run `python ecg-byte_b200/build.py` "
                "(needs nvcc); this package has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        EXPORTS = _declare(L)
        _lib = L
    return _lib


def check(status):
    if status != OK:
        msg = lib().ecgb_last_error().decode("utf-8", "replace")
        if status == EINVAL:
            raise ValueError("libecgbyte: " + msg)
        raise EcgbError(status, msg)


def device_count():
    n = C.c_int(0)
    rc = lib().ecgb_device_count(C.byref(n))
    return n.value if rc == OK else 0


def require_device():
    n = C.c_int(0)
    check(lib().ecgb_device_count(C.byref(n)))
    return n.value
