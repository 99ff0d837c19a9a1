"""Where does the end-to-end path lose against the bare copy?  (a) bare H2D of the whole buffer vs the same bytes in
chunks (one stream / three streams), (b) the two pipelines over chunk size and depth.  1 GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import numpy as np, torch
import bench
from ecgbyte import synth
from ecgbyte.api import EncodePipeline, EncodePipelineCSR, Quantizer, Vocab
dev = torch.device("cuda", 0)
bench.bind_to_gpu_numa(dev)
pairs, pct = bench.load_table()
q = Quantizer(pct, dtype=torch.float32, device=dev)
v = Vocab.from_pairs(pairs, device=dev)
n, stride = 16384, 8192
x = synth.corpus_cuda(2024, n, bench.L_SAMPLES, torch.float32, dev)
xh = torch.empty((n, bench.C_LEADS, bench.L_SAMPLES), dtype=torch.float32).pin_memory(); xh.copy_(x)
d = torch.empty_like(x)
GB = n * bench.REC_LEN * 4 / 1e9

def wall(fn, reps=4):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

print("bare H2D, one copy: %.1f GB/s" % (GB / wall(lambda: d.copy_(xh, non_blocking=True))))
for chunk in (512, 2048, 8192):
    def one_stream():
        for c0 in range(0, n, chunk):
            d[c0:c0 + chunk].copy_(xh[c0:c0 + chunk], non_blocking=True)
    print("bare H2D, chunks of %d on one stream: %.1f GB/s" % (chunk, GB / wall(one_stream)))
    ss = [torch.cuda.Stream(dev) for _ in range(3)]
    def three_streams():
        for k, c0 in enumerate(range(0, n, chunk)):
            with torch.cuda.stream(ss[k % 3]):
                d[c0:c0 + chunk].copy_(xh[c0:c0 + chunk], non_blocking=True)
    print("bare H2D, chunks of %d round-robin on three streams: %.1f GB/s" % (chunk, GB / wall(three_streams)))
tok_h = torch.empty((n, stride), dtype=torch.int32).pin_memory()
tok16 = torch.empty((n * 6144,), dtype=torch.uint16).pin_memory()
len_h = torch.empty((n,), dtype=torch.int32).pin_memory()
for cls, out in ((EncodePipeline, tok_h), (EncodePipelineCSR, tok16)):
    for chunk, depth in ((1024, 3), (2048, 3), (4096, 3), (8192, 3), (2048, 4), (4096, 2)):
        pipe = cls(v, q, bench.REC_LEN, stride, chunk=chunk, depth=depth)
        t = wall(lambda: pipe.run(xh, out, len_h))
        print("%-18s chunk %5d depth %d: %7.2f ms/step  %8.0f records/s  H2D %.1f GB/s" % (cls.__name__, chunk, depth, t * 1e3, n / t, GB / t), flush=True)
        del pipe
