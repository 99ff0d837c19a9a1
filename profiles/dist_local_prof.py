"""Sharded-training kernels on one GPU (2 local 'ranks', config-1 corpus), for an ncu launch list."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import torch
from ecgbyte import synth
from ecgbyte.api import Quantizer
from ecgbyte.dist_train import train_shards_local, split_contiguous
m = int(sys.argv[1]) if len(sys.argv) > 1 else 300
q = Quantizer(synth.BENCH_PERCENTILES, dtype=torch.float32, device="cuda:0")
sym = q.quantize(synth.corpus_cuda(0, 1000, 5000, torch.float32, "cuda:0")).reshape(-1)
shards = [sym[lo:hi] for lo, hi in split_contiguous(sym.numel(), 2)]
torch.cuda.synchronize(); t0 = time.perf_counter()
res, trs = train_shards_local(shards, m)
torch.cuda.synchronize(); print("%d merges, 2 local shards: %.1f ms" % (m, (time.perf_counter() - t0) * 1e3))
