"""Sharded training over real ranks (NCCL):
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/dist_gpu_train.py [records] [merges] [nccl|persistent|auto]
(persistent = one cooperative kernel per rank, device-initiated exchange over NVLink peer memory; nccl = step-wise
launches with two all-gathers per step.)  Every rank trains on its contiguous piece of one corpus string; rank 0 compares the merge
list and the concatenated merged stream with the CPU oracle on the whole string."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ecg-byte_b200")]
import numpy as np
import torch
import torch.distributed as dist


def main():
    n_rec = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    mode = sys.argv[3] if len(sys.argv) > 3 else "persistent"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from ecgbyte import synth
    from ecgbyte.api import Quantizer
    from ecgbyte.dist_train import split_contiguous, train_corpus_auto, train_shard, train_shard_persistent

    x = synth.corpus(5, n_rec, 5000, np.float32)           # same corpus on every rank (seeded)
    q = Quantizer(synth.BENCH_PERCENTILES, dtype=torch.float32, device=local)
    sym = q.quantize(torch.from_numpy(x).cuda()).reshape(-1)
    lo, hi = split_contiguous(sym.numel(), world)[rank]
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    tr = None
    if mode == "auto":
        pairs, counts, ntied = train_corpus_auto(sym[lo:hi].contiguous(), m)
    elif mode == "persistent":
        pairs, counts, ntied, tr = train_shard_persistent(sym[lo:hi].contiguous(), m)
    else:
        pairs, counts, ntied, tr = train_shard(sym[lo:hi].contiguous(), m)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ids = torch.from_numpy(tr.ids().astype(np.int64)).cuda() if tr is not None else torch.zeros(0, dtype=torch.int64, device="cuda")
    lens = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(lens, torch.tensor([ids.numel()], device="cuda"))
    mx = int(max(int(l) for l in lens))
    pad = torch.zeros(mx, dtype=torch.int64, device="cuda")
    pad[: ids.numel()] = ids
    allids = [torch.zeros(mx, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(allids, pad)
    ok = True
    if rank == 0:
        from oracle import oracle as O
        o_ids, o_pairs, o_counts, o_ntied = O.train_pairs(sym.cpu().numpy(), m, fast=True)
        cat = np.concatenate([a[: int(l)].cpu().numpy() for a, l in zip(allids, lens)])
        ok = (np.array_equal(pairs, o_pairs) and np.array_equal(counts, o_counts) and np.array_equal(ntied, o_ntied)
              and (tr is None or np.array_equal(cat, o_ids.astype(np.int64))))
        print("sharded training (" + mode + ") over %d ranks: %d symbols, %d merges in %.3f s (%.0f merges/s) -> %s"
              % (world, sym.numel(), len(pairs), dt, len(pairs) / dt, "PARITY OK" if ok else "PARITY FAILED"), flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
