"""Sweep of the EncodePipeline chunk size / depth for the end-to-end (host buffers) number."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import numpy as np, torch
import bench
from ecgbyte import synth
from ecgbyte.api import EncodePipeline, Quantizer, Vocab
dev = torch.device("cuda", 0)
pairs, pct = bench.load_table()
q = Quantizer(pct, dtype=torch.float32, device=dev)
v = Vocab.from_pairs(pairs, device=dev)
n, stride = 16384, 8192
x = synth.corpus_cuda(2024, n, bench.L_SAMPLES, torch.float32, dev)
xh = torch.empty((n, bench.C_LEADS, bench.L_SAMPLES), dtype=torch.float32).pin_memory(); xh.copy_(x)
tok_h = torch.empty((n, stride), dtype=torch.int32).pin_memory()
len_h = torch.empty((n,), dtype=torch.int32).pin_memory()
for chunk, depth in ((4096, 3), (2048, 3), (1024, 3), (512, 3), (1024, 4), (512, 6), (256, 4)):
    pipe = EncodePipeline(v, q, bench.REC_LEN, stride, chunk=chunk, depth=depth)
    pipe.run(xh, tok_h, len_h); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        pipe.run(xh, tok_h, len_h)
        torch.cuda.synchronize()   # the step's result is on the host
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    print("chunk %5d depth %d: %7.2f ms/step  %8.0f records/s  H2D %.1f GB/s" %
          (chunk, depth, ms, n / ms * 1e3, n * bench.REC_LEN * 4 / ms / 1e6), flush=True)
    del pipe
