"""dist_loop_local_kernel under ncu: the persistent sharded loop with 2 ranks in one cooperative launch on one GPU
(config-1 corpus, 600 merges).  ncu --set full -k regex:dist_loop_local -c 1 python profiles/dist_local_prof2.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import numpy as np, torch
from ecgbyte import synth
from ecgbyte.api import Quantizer
from ecgbyte.dist_train import train_shards_persistent_local
q = Quantizer(synth.BENCH_PERCENTILES, dtype=torch.float32, device="cuda:0")
sym = q.quantize(synth.corpus_cuda(0, 1000, 5000, torch.float32, "cuda:0")).reshape(-1)
n = sym.numel()
res, trs = train_shards_persistent_local([sym[: n // 2].contiguous(), sym[n // 2:].contiguous()], 600, timeout_s=60.0)
print(len(res[0][0]), [t.length() for t in trs])
