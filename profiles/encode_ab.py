"""Times the fused quantise+encode kernel on config-2-shaped data (records x 12 x 5000 fp32, 5,000-merge
table) and checks a sample against the oracle.  A/B: ECGB_ENCODE_V1=1 selects the round-1 bitmap-trie
kernel.  usage: python profiles/encode_ab.py [records] [table: 5000|10000] [dtype: f32|i16|f64]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ecg-byte_b200"))
import numpy as np, torch
from ecgbyte import synth
from ecgbyte.api import Quantizer, Vocab
from oracle import oracle as O

n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
dt = {"f32": torch.float32, "i16": torch.int16, "f64": torch.float64}[sys.argv[3] if len(sys.argv) > 3 else "f32"]
f = np.load(os.path.join(ROOT, "tests", "golden", "ptbxl_1000_m%d.npz" % (10000 if M > 5000 else 5000)))
pairs = f["pairs"].astype(np.uint32)[:M]
pct = {"percentile_1": np.float64(f["pct"][0]), "percentile_99": np.float64(f["pct"][1])}
dev = torch.device("cuda:0")
q = Quantizer(pct, dtype=dt, device=dev)
v = Vocab.from_pairs(pairs, device=dev)
x = synth.corpus_cuda(2024, n_rec, 5000, dt, dev)
stride = 8192
tok = torch.empty((n_rec, stride), dtype=torch.int32, device=dev)
lens = torch.empty((n_rec,), dtype=torch.int32, device=dev)
for _ in range(3):
    v.encode_batch(q, x, out_stride=stride, tokens=tok, lens=lens)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
ev[0].record()
for i in range(5):
    v.encode_batch(q, x, out_stride=stride, tokens=tok, lens=lens)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
T = int(lens.sum().item())
es = {torch.float32: 4, torch.int16: 2, torch.float64: 8}[dt]
alg = n_rec * (60000 * es + 4) + 4 * T
k = float(np.mean(ms))
print("kernel=%s records=%d merges=%d dtype=%s: %.3f ms  %.3f M rec/s  %.1f GB/s  frac=%.3f  info=%s" % (
    "v1" if os.environ.get("ECGB_ENCODE_V1") else "v2", n_rec, M, dt, k, n_rec / k / 1e3, alg / k / 1e6, alg / k / 1e6 / 6553.0, v.info()))
# parity on a sample
idx = np.sort(np.random.default_rng(0).choice(n_rec, size=min(96, n_rec), replace=False))
xs = x[torch.from_numpy(idx).to(dev)].cpu().numpy()
sym = O.quantize(xs, pct["percentile_1"], pct["percentile_99"]).reshape(len(idx), -1)
seq, off = O.expand(pairs)
trie = O.Trie(flat=(seq, off, np.arange(256, 256 + len(pairs), dtype=np.uint32)))
w_tok, w_len = trie.encode_batch(sym, stride)
g_tok = tok[torch.from_numpy(idx).to(dev)].cpu().numpy(); g_len = lens.cpu().numpy()[idx]
ok = np.array_equal(w_len.astype(np.int64), g_len.astype(np.int64)) and all(
    np.array_equal(g_tok[k, : w_len[k]], w_tok[k, : w_len[k]].astype(np.int32)) for k in range(len(idx)))
print("parity vs oracle on %d records: %s" % (len(idx), "OK" if ok else "MISMATCH"))
if not ok:
    for k in range(len(idx)):
        if w_len[k] != g_len[k] or not np.array_equal(g_tok[k, : w_len[k]], w_tok[k, : w_len[k]].astype(np.int32)):
            d = np.flatnonzero(g_tok[k, : min(w_len[k], g_len[k])] != w_tok[k, : min(w_len[k], g_len[k])].astype(np.int32))
            print("record", idx[k], "len", g_len[k], "want", w_len[k], "first diff", d[:3], g_tok[k, d[:3]] if len(d) else None, w_tok[k, d[:3]] if len(d) else None)
            break
    sys.exit(1)
