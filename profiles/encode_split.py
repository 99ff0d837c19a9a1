"""Is the quantiser better OUTSIDE the walkers?  Times (a) quantize_kernel alone, (b) the walker on ready-made u8 symbols
(encode_symbols), (c) both back to back and (d) both overlapped on two streams over quarters of the batch, against the
fused kernel.  usage: python profiles/encode_split.py [records] [5000|10000]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import numpy as np, torch
from ecgbyte import synth
from ecgbyte.api import Quantizer, Vocab

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
f = np.load(os.path.join(ROOT, "tests", "golden", "ptbxl_1000_m%d.npz" % M))
pairs = f["pairs"].astype(np.uint32)
pct = {"percentile_1": np.float64(f["pct"][0]), "percentile_99": np.float64(f["pct"][1])}
dev = torch.device("cuda:0")
q = Quantizer(pct, dtype=torch.float32, device=dev)
v = Vocab.from_pairs(pairs, device=dev)
x = synth.corpus_cuda(2024, n, 5000, torch.float32, dev)
stride = 8192
tok = torch.empty((n, stride), dtype=torch.int32, device=dev)
lens = torch.empty((n,), dtype=torch.int32, device=dev)
sym = torch.empty((n, 60000), dtype=torch.uint8, device=dev)

def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

t_fused = timeit(lambda: v.encode_batch(q, x, out_stride=stride, tokens=tok, lens=lens))
ref_lens = lens.clone(); ref_tok = tok[:64].clone()
t_q = timeit(lambda: q.quantize(x, out=sym))
t_w = timeit(lambda: v.encode_symbols(sym, out_stride=stride, tokens=tok, lens=lens))
assert torch.equal(lens, ref_lens) and torch.equal(tok[:64], ref_tok)
t_seq = timeit(lambda: (q.quantize(x, out=sym), v.encode_symbols(sym, out_stride=stride, tokens=tok, lens=lens)))
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def overlapped(parts):
    cur = torch.cuda.current_stream(dev)
    s1.wait_stream(cur); s2.wait_stream(cur)
    step = (n + parts - 1) // parts
    evs = []
    for p in range(parts):
        a, b = p * step, min(n, (p + 1) * step)
        with torch.cuda.stream(s1):
            q.quantize(x[a:b], out=sym[a:b])
            ev = torch.cuda.Event(); ev.record(); evs.append(ev)
        with torch.cuda.stream(s2):
            s2.wait_event(ev)
            v.encode_symbols(sym[a:b], out_stride=stride, tokens=tok[a:b], lens=lens[a:b])
    cur.wait_stream(s1); cur.wait_stream(s2)
res = {}
for parts in (2, 4, 8):
    res[parts] = timeit(lambda: overlapped(parts))
    assert torch.equal(lens, ref_lens) and torch.equal(tok[:64], ref_tok)
print("records %d, %d merges: fused %.2f ms | quantize %.2f | walker on u8 symbols %.2f | back to back %.2f | overlapped %s"
      % (n, M, t_fused, t_q, t_w, t_seq, "  ".join("%d parts %.2f" % (k, r) for k, r in res.items())))
