"""One config-1-scale training run (used under ncu by profiles/capture.sh)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import numpy as np
import torch
from ecgbyte import synth
from ecgbyte.api import Quantizer, Trainer

x = synth.corpus_cuda(0, 1000, 5000, torch.float32, "cuda:0")
q = Quantizer(synth.BENCH_PERCENTILES, dtype=torch.float32, device="cuda:0")
sym = q.quantize(x).reshape(-1)
tr = Trainer(sym.numel(), 3000, device="cuda:0")
tr.load(sym)
pairs, counts, ntied = tr.run(3000)
print(len(pairs), int(tr.lengths(3000)[-1]))
