import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/ecg-byte_b200')
import numpy as np
from ecgbyte.api import Trainer
from oracle import oracle as O
rng = np.random.default_rng(17)
text = rng.integers(97, 101, size=17).astype(np.uint8)
print(bytes(text))
tr = Trainer(len(text), 10); tr.load(text.tobytes())
pairs, counts, ntied = tr.run(10)
print('gpu', pairs.tolist(), counts.tolist(), ntied.tolist(), tr.ids().tolist())
o = O.train_pairs(text, 10)
print('cpu', o[1].tolist(), o[2].tolist(), o[3].tolist(), o[0].tolist())
