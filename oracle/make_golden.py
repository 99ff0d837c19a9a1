"""Generates tests/golden/*.npz by running the REFERENCE's own Python code.

Run in the build container only (needs /root/reference):  python oracle/make_golden.py
The reference's ecg_byte/utils/tokenizer_utils.py is imported unmodified, with
sys.modules stubs for the three modules it imports but that are absent here
(matplotlib, matplotlib.pyplot -- plotting only -- and rust_bpe, the native crate that
cannot be built without a Rust toolchain).  Its normalize_all / process_ecg are then
executed on seeded inputs and the outputs stored next to the inputs.

NumPy here is 2.3.x (the reference pins 1.26.3): float32 input is promoted to float64
by the np.float64 percentiles, which is the float64 contract of SURVEY.md 8a Q1.
"""
import os
import sys
import types
import warnings

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference_tu():
    for name in ("matplotlib", "matplotlib.pyplot", "rust_bpe"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import ecg_byte.utils.tokenizer_utils as tu
    return tu


def main():
    tu = import_reference_tu()
    os.makedirs(OUT, exist_ok=True)
    sys.path.insert(0, os.path.join(os.path.dirname(OUT), "..", "ecg-byte_b200"))
    from ecgbyte import synth

    rng = np.random.default_rng(20240229)
    cases = {}
    # (a) synthetic ECG records with their own stats dict, float64 (the reference's stored type)
    x = synth.corpus(3, 4, L=500, dtype=np.float64)
    pct = synth.percentiles(x, seed=3)
    cases["ecg_f64"] = (x, pct["percentile_1"], pct["percentile_99"])
    # (b) the same records stored as float32
    cases["ecg_f32"] = (x.astype(np.float32), pct["percentile_1"], pct["percentile_99"])
    # (c) dense sweep across all 26 bins plus out-of-range values and specials
    p1, p99 = np.float64(-0.37), np.float64(1.21)
    sweep = np.linspace(p1 - 1.5, p99 + 1.5, 20001)
    special = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1e300, -1e300, 5e-324])
    cases["sweep_f64"] = (np.concatenate([sweep, special]), p1, p99)
    cases["sweep_f32"] = (np.concatenate([sweep, special]).astype(np.float32), p1, p99)
    # (d) values straddling every bin edge by a few ulps (edges located with the reference itself)
    lo, den = (p1 - 0.5), ((p99 + 0.5) - (p1 - 0.5) + 1e-6)
    edges = lo + den * np.arange(1, 26) / 26.0
    near = np.concatenate([np.nextafter(edges, -np.inf), edges, np.nextafter(edges, np.inf)])
    for _ in range(3):
        near = np.concatenate([near, np.nextafter(near, -np.inf), np.nextafter(near, np.inf)])
    cases["edges_f64"] = (near, p1, p99)
    e32 = edges.astype(np.float32)
    near32 = np.concatenate([np.nextafter(e32, np.float32(-np.inf)), e32, np.nextafter(e32, np.float32(np.inf))])
    for _ in range(3):
        near32 = np.concatenate([near32, np.nextafter(near32, np.float32(-np.inf)), np.nextafter(near32, np.float32(np.inf))])
    cases["edges_f32"] = (near32, p1, p99)
    # (e) random percentiles / random data
    for i in range(4):
        a = np.float64(rng.normal(0, 2))
        b = np.float64(a + abs(rng.normal(0, 3)))
        cases["rand%d_f64" % i] = (rng.normal((a + b) / 2, (b - a + 1), size=5000), a, b)
        cases["rand%d_f32" % i] = (rng.normal((a + b) / 2, (b - a + 1), size=5000).astype(np.float32), a, b)

    store = {}
    for name, (sig, a, b) in cases.items():
        pcts = {"percentile_1": np.float64(a), "percentile_99": np.float64(b)}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            clipped, symbols = tu.normalize_all(sig, pcts)      # the reference, unmodified
        codes = np.char.encode(symbols.reshape(-1), "ascii").view(np.uint8).reshape(sig.shape)
        store[name + "__in"] = sig
        store[name + "__pct"] = np.array([a, b], np.float64)
        store[name + "__sym"] = codes
        if name == "ecg_f64":
            store[name + "__clipped"] = clipped
            # process_ecg's string form: ''.join(symbol_signal.flatten()) per record (tu.py:59)
            store[name + "__str0"] = np.frombuffer("".join(symbols[0].flatten()).encode(), np.uint8)
    np.savez_compressed(os.path.join(OUT, "quantize_reference.npz"), **store)
    print("wrote", os.path.join(OUT, "quantize_reference.npz"), len(cases), "cases; numpy", np.__version__)


if __name__ == "__main__":
    main()
