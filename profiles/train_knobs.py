"""A/B of trainer tuning knobs inside one process (same box, same clocks).
usage: train_knobs.py ENV_NAME v1,v2,...  [n_records] [merges]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import torch
from ecgbyte import synth
from ecgbyte.api import Quantizer, Trainer
name, values = sys.argv[1], sys.argv[2].split(",")
n_rec = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
merges = int(sys.argv[4]) if len(sys.argv) > 4 else 5000
q = Quantizer(synth.BENCH_PERCENTILES, dtype=torch.float32, device="cuda:0")
sym = q.quantize(synth.corpus_cuda(0, n_rec, 5000, torch.float32, "cuda:0")).reshape(-1)
tr = Trainer(sym.numel(), merges, device="cuda:0")
for rep in range(2):
    for val in values:
        os.environ[name] = val
        best = 1e9
        for _ in range(2):
            tr.load(sym); torch.cuda.synchronize()
            t0 = time.perf_counter(); tr.run(merges); best = min(best, time.perf_counter() - t0)
        print("%s=%-8s records %d merges %d: %8.2f ms" % (name, val, n_rec, merges, best * 1e3), flush=True)
