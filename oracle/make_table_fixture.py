"""Generates tests/golden/ptbxl_1000_m5000.npz with the ORACLE (test infrastructure):
the BASELINE.json config-1 corpus (1,000 synthetic PTB-XL-shaped records, seed 0), its
stats dict, and the M-merge table learned by oracle.ecgo_train_fast (~75 s here for M = 5,000).
Run:  python oracle/make_table_fixture.py [M]      (M = 5000: config 2's table; M = 10000: config 3's)"""
import sys, time
sys.path.insert(0, '/root/repo/ecg-byte_b200'); sys.path.insert(0, '/root/repo')
import numpy as np
M = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
from ecgbyte import synth
from oracle import oracle as O
t = time.time()
X = synth.corpus(0, 1000, 5000, np.float32)
pct = synth.percentiles(X, seed=0)
print('gen', time.time() - t, pct)
S = O.quantize(X, pct['percentile_1'], pct['percentile_99'])
t = time.time()
ids, pairs, counts, ntied = O.train_pairs(S.reshape(-1), M, fast=True)
print('train', time.time() - t, 'len', len(ids), 'compression', S.size / len(ids), 'ties', int((ntied > 1).sum()))
np.savez_compressed('/root/repo/tests/golden/ptbxl_1000_m%d.npz' % M, pairs=pairs.astype(np.uint16), counts=counts, ntied=ntied,
                    pct=np.array([pct['percentile_1'], pct['percentile_99']]), n_ids=np.array([len(ids)]),
                    ids_crc=np.array([int(np.bitwise_xor.reduce(ids.astype(np.uint64) * (np.arange(len(ids), dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1))))], np.uint64))
