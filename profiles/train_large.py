"""Streaming regime of the trainer: a 1.2e9-symbol corpus (20k records), a few hundred merges.
Reports the achieved 2*(n_t + n_{t+1}) bytes/s over step ranges (HBM-bound regime, unlike config 1's tail)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import numpy as np, torch
from ecgbyte import synth
from ecgbyte.api import Quantizer, Trainer
n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
x = synth.corpus_cuda(0, n_rec, 5000, torch.float32, "cuda:0")
q = Quantizer(synth.BENCH_PERCENTILES, dtype=torch.float32, device="cuda:0")
sym = q.quantize(x).reshape(-1)
del x
tr = Trainer(sym.numel(), 1000, device="cuda:0")
prev_t, prev_m = 0.0, 0
for m in (10, 50, 200, 1000):
    tr.load(sym); torch.cuda.synchronize()
    t0 = time.perf_counter(); tr.run(m); dt = time.perf_counter() - t0
    n = tr.lengths(m).astype(np.float64)
    seg = np.sum(2.0 * (n[prev_m:m] + n[prev_m + 1:m + 1]))
    print("steps %4d..%4d: %.1f ms, n_t %.2e -> %.2e, %.0f GB/s of the 2(n_t+n_t+1) formula" %
          (prev_m, m, (dt - prev_t) * 1e3, n[prev_m], n[m], seg / (dt - prev_t) / 1e9), flush=True)
    prev_t, prev_m = dt, m
