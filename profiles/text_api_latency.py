"""Latency of the literal drop-in call rust_bpe.encode_text(text, merges) (one string per call, host str in,
Python list out) next to the CPU port of the reference (trie rebuilt per call, lib.rs:153-161)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import numpy as np
import rust_bpe
from ecgbyte import synth
from ecgbyte.api import expand_merges
from oracle import oracle as O

f = np.load(os.path.join(ROOT, "tests", "golden", "ptbxl_1000_m5000.npz"))
pairs = f["pairs"].astype(np.uint32)
seq, off = expand_merges(pairs)
merges = [(seq[int(off[i]):int(off[i + 1])].tolist(), 256 + i) for i in range(len(pairs))]
x = synth.corpus(5, 64, 5000, np.float32)
sym = O.quantize(x, f["pct"][0], f["pct"][1]).reshape(64, -1)
for n in (6000, 60000, 64 * 60000):
    s = sym.reshape(-1)[:n].tobytes().decode()
    got = rust_bpe.encode_text(s, merges)
    reps = 20 if n <= 60000 else 3
    t0 = time.perf_counter()
    for _ in range(reps):
        got = rust_bpe.encode_text(s, merges)
    t_gpu = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(max(reps // 4, 1)):
        want = O.encode_text(s, merges)
    t_cpu = (time.perf_counter() - t0) / max(reps // 4, 1)
    assert got == want
    print("encode_text, %8d symbols: GPU %.3f ms per call, CPU port %.3f ms per call (%d tokens)" % (n, t_gpu * 1e3, t_cpu * 1e3, len(got)), flush=True)
