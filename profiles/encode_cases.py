"""Both encode kernels on the shapes that matter besides the headline batch (run once with ECGB_ENCODE_V1=1, once with
ECGB_ENCODE_V2=1): small data-loader batches, an e2e chunk, other sample types.  Prints ms per call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import numpy as np, torch
from ecgbyte import synth
from ecgbyte.api import Quantizer, Vocab
f = np.load(os.path.join(ROOT, "tests", "golden", "ptbxl_1000_m5000.npz"))
pairs = f["pairs"].astype(np.uint32)
pct = {"percentile_1": np.float64(f["pct"][0]), "percentile_99": np.float64(f["pct"][1])}
dev = torch.device("cuda:0")
v = Vocab.from_pairs(pairs, device=dev)
which = "v1" if os.environ.get("ECGB_ENCODE_V1") else ("v2" if os.environ.get("ECGB_ENCODE_V2") else "auto")
out = []
for n, L, dt in ((2, 500, torch.float32), (64, 500, torch.float64), (2048, 5000, torch.float32), (16384, 5000, torch.float32),
                 (100000, 5000, torch.int16), (50000, 5000, torch.float64), (100000, 2500, torch.float32)):
    q = Quantizer(pct, dtype=dt, device=dev)
    x = synth.corpus_cuda(7, n, L, dt, dev)
    stride = 12 * L // 6
    tok = torch.empty((n, stride), dtype=torch.int32, device=dev)
    lens = torch.empty((n,), dtype=torch.int32, device=dev)
    for _ in range(3):
        v.encode_batch(q, x, out_stride=stride, tokens=tok, lens=lens)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20 if n <= 2048 else 5
    e0.record()
    for _ in range(reps):
        v.encode_batch(q, x, out_stride=stride, tokens=tok, lens=lens)
    e1.record(); torch.cuda.synchronize()
    out.append("%dx12x%d %s: %.3f ms (sum len %d)" % (n, L, str(dt).split(".")[1], e0.elapsed_time(e1) / reps, int(lens.sum())))
    del x, tok
print(which, " | ".join(out))
