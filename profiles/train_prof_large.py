import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import torch
from ecgbyte import synth
from ecgbyte.api import Quantizer, Trainer
x = synth.corpus_cuda(0, 20000, 5000, torch.float32, "cuda:0")
q = Quantizer(synth.BENCH_PERCENTILES, dtype=torch.float32, device="cuda:0")
sym = q.quantize(x).reshape(-1)
del x
tr = Trainer(sym.numel(), 100, device="cuda:0")
tr.load(sym)
print(len(tr.run(100)[0]), tr.length())
