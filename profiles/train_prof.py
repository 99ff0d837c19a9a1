"""Short training run for ncu (argv[1] = number of merges)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import torch
from ecgbyte import synth
from ecgbyte.api import Quantizer, Trainer
m = int(sys.argv[1]) if len(sys.argv) > 1 else 300
x = synth.corpus_cuda(0, 1000, 5000, torch.float32, "cuda:0")
q = Quantizer(synth.BENCH_PERCENTILES, dtype=torch.float32, device="cuda:0")
sym = q.quantize(x).reshape(-1)
tr = Trainer(sym.numel(), m, device="cuda:0")
tr.load(sym)
print(len(tr.run(m)[0]), tr.length())
