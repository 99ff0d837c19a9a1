"""Phase breakdown of train_loop_kernel for one CTA (needs a build with ECGB_NVCC_EXTRA=-DECGB_TRAIN_TIMING).
The library prints the accumulated phase times of CTA 1 / thread 0 to stderr after each run."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import torch
from ecgbyte import synth
from ecgbyte.api import Quantizer, Trainer
n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
merges = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
q = Quantizer(synth.BENCH_PERCENTILES, dtype=torch.float32, device="cuda:0")
sym = q.quantize(synth.corpus_cuda(0, n_rec, 5000, torch.float32, "cuda:0")).reshape(-1)
tr = Trainer(sym.numel(), merges, device="cuda:0")
for m in (1000, merges):
    tr.load(sym); torch.cuda.synchronize()
    t0 = time.perf_counter(); tr.run(m); dt = time.perf_counter() - t0
    sys.stderr.flush()
    print("records %d merges %d: %.2f ms" % (n_rec, m, dt * 1e3), flush=True)
