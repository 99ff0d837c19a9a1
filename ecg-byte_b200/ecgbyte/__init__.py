"""ecgbyte -- B200-native ECG-Byte tokenizer hot path (quantise, encode, train).

Host-side mirror of the reference interfaces for this path:
  ecgbyte.tokenizer_utils  <-> ecg_byte/utils/tokenizer_utils.py
  rust_bpe (sibling module) <-> the PyO3 module built from ecg_byte/rust_bpe/src/lib.rs
All compute runs in libecgbyte.so (CUDA, sm_100a); there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import EcgbError, device_count  # noqa: F401

__all__ = ["Quantizer", "Vocab", "Trainer", "EcgbError", "device_count"]


def __getattr__(name):
    # torch is imported lazily so that `import ecgbyte` stays cheap for tooling
    if name in ("Quantizer", "Vocab", "Trainer", "expand_merges", "flatten_merges"):
        from . import api
        return getattr(api, name)
    raise AttributeError(name)
