"""BASELINE config 4 (and scaled-down versions of it): BPE training over a corpus of `records` PTB-XL-shaped records.
    python profiles/train_config4.py [records] [merges] [table_log2]                       (one GPU: train_loop_kernel)
    torchrun --nproc-per-node N profiles/train_config4.py [records] [merges] [table_log2]  (N GPUs: dist_loop_kernel)
Reports seconds, merges/s and the fraction of the HBM roofline on the 2(n_t + n_{t+1}) formula."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import numpy as np, torch

n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
tlog = int(sys.argv[3]) if len(sys.argv) > 3 else 24
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
from ecgbyte import synth
from ecgbyte.api import Quantizer, Trainer
from ecgbyte.dist_train import ShardedTrainer, split_contiguous

q = Quantizer(synth.BENCH_PERCENTILES, dtype=torch.float32, device=dev)
lo, hi = split_contiguous(n_rec, world)[rank]
t0 = time.perf_counter()
shard = torch.empty((hi - lo) * 60000, dtype=torch.uint8, device=dev)
for a in range(lo, hi, 2048):
    b = min(hi, a + 2048)
    shard[(a - lo) * 60000:(b - lo) * 60000] = q.quantize(synth.corpus_cuda_range(0, n_rec, a, b, 5000, torch.float32, dev)).reshape(-1)
torch.cuda.synchronize()
gen = time.perf_counter() - t0
if world == 1:
    tr = Trainer(shard.numel(), m, device=dev, table_log2=tlog)
    tr.load(shard)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pairs, counts, ntied = tr.run(m)
    dt = time.perf_counter() - t0
    lens = tr.lengths(len(pairs)).astype(np.float64)
    print("pair table:", tr.table_stats(), file=sys.stderr, flush=True)
else:
    st = ShardedTrainer(shard.numel(), m, table_log2=tlog)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    pairs, counts, ntied = st.train(shard, m)
    d = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(d, op=dist.ReduceOp.MAX)
    dt = float(d)
    l = torch.from_numpy(st.tr.lengths(len(pairs)).astype(np.float64)).to(dev)
    dist.all_reduce(l)
    lens = l.cpu().numpy()
    if rank == 0:
        print("pair table:", st.tr.table_stats(), file=sys.stderr, flush=True)
alg = float(np.sum(2.0 * (lens[:-1] + lens[1:])))
if rank == 0:
    peak = 6553.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    print(json.dumps({"records": n_rec, "symbols": int(lens[0]), "merges": int(len(pairs)), "gpus": world, "seconds": dt,
                      "merges_per_s": len(pairs) / dt, "final_tokens": int(lens[-1]), "algorithmic_bytes": alg,
                      "achieved_gbs": alg / dt / 1e9, "frac_of_hbm_peak_per_gpu": alg / dt / 1e9 / peak / world,
                      "generate_s": gen, "first_pairs": pairs[:4].tolist(), "last_pair": pairs[-1].tolist(),
                      "crc": int(np.bitwise_xor.reduce(pairs.astype(np.uint64).reshape(-1) * np.arange(1, 2 * len(pairs) + 1, dtype=np.uint64)))}),
          flush=True)
if world > 1:
    dist.barrier()
    st.close()
    dist.destroy_process_group()
