"""BASELINE.json config 5: on-the-fly tokenisation in front of a random-init Llama-3.2-1B-shaped
LLM (batch 2, pad_to_max 1020, records of 12 x 500 samples as in the reference's scripts).
Prints the batcher latency (quantise + encode + pack, host lists in -> device tensors out) next to the
LLM's forward+backward step time.  The LLM is plain HuggingFace code and is not part of this repo."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import numpy as np, torch
from ecgbyte import synth
from ecgbyte.api import expand_merges
from ecgbyte.data_loader import ECGTokenBatcher

f = np.load(os.path.join(ROOT, "tests", "golden", "ptbxl_1000_m5000.npz"))
pairs = f["pairs"].astype(np.uint32)
pct = {"percentile_1": f["pct"][0], "percentile_99": f["pct"][1]}
seq, off = expand_merges(pairs)
merges = [(seq[int(off[i]):int(off[i + 1])].tolist(), 256 + i) for i in range(len(pairs))]
V = 128256 + 3 + 256 + len(pairs)
lut = torch.arange(256 + len(pairs), dtype=torch.int64) + 128259
rng = np.random.default_rng(0)
for bs in (2, 64, 1024):
    b = ECGTokenBatcher(merges, pct, lut, pad_to_max=1020, pad_id=128256, bos_id=128000, eos_id=128001,
                        sig_start_id=128257, sig_end_id=128258, dtype=torch.float64)
    x = synth.corpus(3, bs, L=500, dtype=np.float64)
    qs = [rng.integers(0, 128000, size=12).tolist() for _ in range(bs)]
    ans = [rng.integers(0, 128000, size=40).tolist() for _ in range(bs)]
    for _ in range(3):
        out = b(x, qs, ans)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        out = b(x, qs, ans)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print("batch %5d: tokenise+pack %.3f ms per batch (%.0f records/s), row %s" % (bs, dt * 1e3, bs / dt, tuple(out["tokenized_signal"].shape)), flush=True)
try:
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(vocab_size=V, hidden_size=2048, intermediate_size=8192, num_hidden_layers=16, num_attention_heads=32,
                      num_key_value_heads=8, max_position_embeddings=2048, rms_norm_eps=1e-5, rope_theta=500000.0)
    model = LlamaForCausalLM(cfg).to(torch.bfloat16).cuda()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5)
    ids = out["tokenized_signal"][:2].clone()
    labels = out["quantized_signal_ids_input"][:2].clone()
    attn = out["attn_mask"][:2].clone()
    def step():
        loss = model(input_ids=ids, attention_mask=attn, labels=labels).loss
        loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
        return loss
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print("Llama-3.2-1B-shaped random init, batch 2 x 1024, bf16, full fine-tune step: %.1f ms" % (dt * 1e3))
except Exception as e:
    print("LLM step not measured:", repr(e)[:200])
