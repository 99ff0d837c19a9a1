"""Wall time of the training loop as a function of the number of merges (where does the time go?)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import numpy as np, torch
from ecgbyte import synth
from ecgbyte.api import Quantizer, Trainer
x = synth.corpus_cuda(0, 1000, 5000, torch.float32, "cuda:0")
q = Quantizer(synth.BENCH_PERCENTILES, dtype=torch.float32, device="cuda:0")
sym = q.quantize(x).reshape(-1)
for log2 in (22, 20):
    tr = Trainer(sym.numel(), 5000, device="cuda:0", table_log2=log2)
    for m in (0, 1, 10, 50, 200, 1000, 5000):
        best = 1e9
        for _ in range(2):
            tr.load(sym); torch.cuda.synchronize()
            t0 = time.perf_counter(); tr.run(m); dt = time.perf_counter() - t0
            best = min(best, dt)
        print("table 2^%d  merges %5d  %8.2f ms  tokens left %d" % (log2, m, best * 1e3, tr.length()), flush=True)
