"""Sharded byte_pair_encoding (lib.rs:58-125) over the GPUs of one box.

The reference trains on ONE string (tokenizer_utils.py:93), so pairs are counted and
merged across record boundaries.  Here rank g owns a contiguous piece of that string
(rank order == string order).  Per merge step every rank:

  commit : applies every rank's histogram-delta list to ITS copy of the global pair
           histogram (all copies stay identical, so the argmax with the deterministic tie
           rule is identical everywhere -- no reduction, no candidate certification),
           takes the argmax, publishes a 64-byte boundary record (shard length, first 3 /
           last 2 tokens, parity of its trailing run of `left` for (x,x) merges);
  merge  : merges the winning pair in its shard.  The 2-token left halo, 3-token right
           halo and the parity of the (x,x) run entering the shard are derived from the
           gathered boundary records; a pair that straddles two shards belongs to the left
           one (the right one drops its first token).  The histogram patches go to a
           delta list.

The two exchanges per step are all-gathers of small fixed-size device buffers
(NCCL over NVLink through torch.distributed); everything else is on-device and
asynchronous.  With world == 1 this degenerates to the single-device loop.
"""
import numpy as np
import torch

from .api import Trainer


class TorchExchange:
    """all-gather through torch.distributed (NCCL on GPUs)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_gather(self, out, inp):
        self.dist.all_gather_into_tensor(out, inp, group=self.group)


def train_shard(shard, num_merges, exchange=None, device=None, table_log2=0, check_every=0):
    """shard: this rank's piece of the corpus (bytes / numpy uint8 / uint8 CUDA tensor).
    Returns (pairs [m,2], counts [m], ntied [m], trainer); identical on every rank."""
    ex = exchange or TorchExchange()
    n = shard.numel() if isinstance(shard, torch.Tensor) else len(shard)
    tr = Trainer(max(n, 1), num_merges, device=device, table_log2=table_log2)
    tr.load(shard)
    dev = torch.device("cuda", tr.device)
    bbytes, lbytes = tr.dist_sizes()
    bnd = torch.zeros(bbytes, dtype=torch.uint8, device=dev)
    all_bnd = torch.zeros(bbytes * ex.world, dtype=torch.uint8, device=dev)
    lst = torch.zeros(lbytes, dtype=torch.uint8, device=dev)
    all_lst = torch.zeros(lbytes * ex.world, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        tr.dist_begin(ex.rank, ex.world, bnd)
        ex.all_gather(all_bnd, bnd)
        tr.dist_count(all_bnd, lst)
        ex.all_gather(all_lst, lst)
        for step in range(num_merges):
            tr.dist_commit(step, all_lst, bnd)
            ex.all_gather(all_bnd, bnd)
            tr.dist_merge(step, all_bnd, lst)
            ex.all_gather(all_lst, lst)
        pairs, counts, ntied = tr.results(num_merges)
    return pairs, counts, ntied, tr


def train_shards_local(shards, num_merges, device=None, table_log2=0):
    """The same protocol with every 'rank' living in this process on ONE device (the
    all-gathers become concatenations).  Used to test the sharded kernels on a single GPU."""
    world = len(shards)
    trs = []
    for s in shards:
        n = s.numel() if isinstance(s, torch.Tensor) else len(s)
        t = Trainer(max(n, 1), num_merges, device=device, table_log2=table_log2)
        t.load(s)
        trs.append(t)
    dev = torch.device("cuda", trs[0].device)
    bbytes, lbytes = trs[0].dist_sizes()
    all_bnd = torch.zeros(bbytes * world, dtype=torch.uint8, device=dev)
    all_lst = torch.zeros(lbytes * world, dtype=torch.uint8, device=dev)
    bnd = [all_bnd[r * bbytes:(r + 1) * bbytes] for r in range(world)]
    lst = [all_lst[r * lbytes:(r + 1) * lbytes] for r in range(world)]
    # separate staging buffers: a rank must not overwrite its slot while others still read it
    nbnd = torch.zeros_like(all_bnd)
    nlst = torch.zeros_like(all_lst)
    with torch.cuda.device(dev):
        for r, t in enumerate(trs):
            t.dist_begin(r, world, nbnd[r * bbytes:(r + 1) * bbytes])
        all_bnd.copy_(nbnd)
        for r, t in enumerate(trs):
            t.dist_count(all_bnd, nlst[r * lbytes:(r + 1) * lbytes])
        all_lst.copy_(nlst)
        for step in range(num_merges):
            for r, t in enumerate(trs):
                t.dist_commit(step, all_lst, nbnd[r * bbytes:(r + 1) * bbytes])
            all_bnd.copy_(nbnd)
            for r, t in enumerate(trs):
                t.dist_merge(step, all_bnd, nlst[r * lbytes:(r + 1) * lbytes])
            all_lst.copy_(nlst)
        res = [t.results(num_merges) for t in trs]
    del bnd, lst
    return res, trs


def split_contiguous(n, world):
    """[lo, hi) of each rank's piece of an n-byte corpus."""
    return [(n * r // world, n * (r + 1) // world) for r in range(world)]
