"""torchrun --nproc-per-node N profiles/dist_persistent.py [records] [merges] [reps]
Times the persistent sharded trainer (ShardedTrainer.train: load + get_stats exchange + dist_loop_kernel) over N ranks,
set-up (allocation, IPC mapping) excluded; rank 0 checks the merge list against single-GPU training of the whole corpus."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import numpy as np, torch, torch.distributed as dist

n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from ecgbyte import synth
from ecgbyte.api import Quantizer, Trainer
from ecgbyte.dist_train import ShardedTrainer, split_contiguous

dev = torch.device("cuda", local)
q = Quantizer(synth.BENCH_PERCENTILES, dtype=torch.float32, device=dev)
lo_r, hi_r = split_contiguous(n_rec, world)[rank]      # whole records per rank (contiguous in corpus order)
parts = []
for a in range(lo_r, hi_r, 2048):
    b = min(hi_r, a + 2048)
    parts.append(q.quantize(synth.corpus_cuda_range(0, n_rec, a, b, 5000, torch.float32, dev)).reshape(-1))
shard = torch.cat(parts) if parts else torch.zeros(0, dtype=torch.uint8, device=dev)
del parts
t0 = time.perf_counter()
st = ShardedTrainer(shard.numel(), m)
torch.cuda.synchronize(); dist.barrier()
setup = time.perf_counter() - t0
best = None
for _ in range(reps):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    pairs, counts, ntied = st.train(shard, m)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    best = float(dt) if best is None else min(best, float(dt))
n_total = torch.tensor([shard.numel()], dtype=torch.int64, device=dev)
dist.all_reduce(n_total)
msg = ""
if rank == 0 and n_rec <= 4000 and world > 1:
    # reference: the whole corpus on this one GPU
    full = q.quantize(synth.corpus_cuda(0, n_rec, 5000, torch.float32, dev)).reshape(-1)
    tr = Trainer(full.numel(), m, device=dev)
    tr.load(full)
    t0 = time.perf_counter()
    p1, c1, t1 = tr.run(m)
    single = time.perf_counter() - t0
    ok = np.array_equal(p1, pairs) and np.array_equal(c1, counts) and np.array_equal(t1, ntied)
    msg = " | single GPU %.3f s (%.0f merges/s) | merge list %s" % (single, len(p1) / single, "EQUAL" if ok else "DIFFERS")
if rank == 0:
    print("persistent sharded training x%d: %d symbols, %d merges: %.3f s (%.0f merges/s), set-up %.2f s%s"
          % (world, int(n_total), len(pairs), best, len(pairs) / best, setup, msg), flush=True)
dist.barrier()
st.close()
dist.destroy_process_group()
