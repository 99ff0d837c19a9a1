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

The step number lives in a counter on the device (Trainer.STEP_DEVICE), so every step issues
the same calls with the same arguments and one step can be captured in a CUDA graph (kernels
and both all-gathers) and replayed for the remaining merges -- one graph launch per merge
instead of ~8 launches and two Python-level collectives.  Measured on 2 B200s the step is bound
by the device (kernels ~45 us + two all-gathers), not by the host: 13.1 k merges/s replayed vs
12.9 k eager, so the NCCL path stays eager unless ECGB_DIST_GRAPH=1; the single-process variant
(train_shards_local), where the Python loop over the ranks does dominate, replays by default.
"""
import os

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

    def all_gather_object(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def barrier(self):
        self.dist.barrier(group=self.group)


def _graph_default(default):
    v = os.environ.get("ECGB_DIST_GRAPH")
    return default if v is None else v != "0"


def train_shard(shard, num_merges, exchange=None, device=None, table_log2=0, check_every=0, use_graph=None):
    """shard: this rank's piece of the corpus (bytes / numpy uint8 / uint8 CUDA tensor).
    Returns (pairs [m,2], counts [m], ntied [m], trainer); identical on every rank."""
    ex = exchange or TorchExchange()
    if use_graph is None:
        use_graph = _graph_default(False)
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
        S = Trainer.STEP_DEVICE

        def one_step():
            tr.dist_commit(S, all_lst, bnd)
            ex.all_gather(all_bnd, bnd)
            tr.dist_merge(S, all_bnd, lst)
            ex.all_gather(all_lst, lst)
            tr.dist_advance(dev)

        _run_steps(one_step, num_merges, dev, use_graph)
        pairs, counts, ntied = tr.results(num_merges)
    return pairs, counts, ntied, tr


class ShardedTrainer:
    """The persistent sharded loop (csrc/train.cu dist_loop_kernel), one process per GPU: ONE cooperative kernel
    per rank runs every merge step and exchanges histogram patches and shard records with the peers by writing
    straight into their memory over NVLink (CUDA IPC mappings of each rank's receive area).  torch.distributed is
    used for the set-up only: IPC handles, the initial boundary records and get_stats lists, two barriers."""

    def __init__(self, capacity_tokens, max_merges, exchange=None, device=None, table_log2=0):
        import ctypes as C
        from . import _lib
        self.ex = exchange or TorchExchange()
        self.tr = Trainer(max(int(capacity_tokens), 1), max_merges, device=device, table_log2=table_log2)
        self.dev = torch.device("cuda", self.tr.device)
        ex, tr = self.ex, self.tr
        self.area, self.area_bytes = tr.peer_area(ex.world)
        h = (C.c_ubyte * 64)()
        _lib.check(_lib.lib().ecgb_ipc_export(C.c_void_p(self.area), h))
        handles = ex.all_gather_object(bytes(h))
        self.areas, self._opened = [], []
        for r in range(ex.world):
            if r == ex.rank:
                self.areas.append(self.area)
                continue
            p = C.c_void_p()
            _lib.check(_lib.lib().ecgb_ipc_open(handles[r], tr.device, C.byref(p)))
            self.areas.append(p.value)
            self._opened.append(p.value)
        bbytes, lbytes = tr.dist_sizes()
        self._bnd = torch.zeros(bbytes, dtype=torch.uint8, device=self.dev)
        self._all_bnd = torch.zeros(bbytes * ex.world, dtype=torch.uint8, device=self.dev)
        self._lst = torch.zeros(lbytes, dtype=torch.uint8, device=self.dev)
        self._all_lst = torch.zeros(lbytes * ex.world, dtype=torch.uint8, device=self.dev)

    def train(self, shard, num_merges, max_ctas=0, timeout_s=0.0):
        """shard: this rank's contiguous piece of the corpus.  -> (pairs, counts, ntied), identical on every rank."""
        ex, tr = self.ex, self.tr
        tr.load(shard)
        with torch.cuda.device(self.dev):
            tr.peer_area(ex.world)  # clears headers and counters; the collectives below order it before any peer's run
            tr.dist_begin(ex.rank, ex.world, self._bnd)
            ex.all_gather(self._all_bnd, self._bnd)
            tr.dist_count(self._all_bnd, self._lst)
            ex.all_gather(self._all_lst, self._lst)
            tr.dist_apply(self._all_lst, ex.world)
            torch.cuda.synchronize(self.dev)
            ex.barrier()
            tr.dist_run(ex.rank, ex.world, self.areas, self._all_bnd, num_merges, max_ctas=max_ctas, timeout_s=timeout_s)
            return tr.results(num_merges)

    def close(self):
        from . import _lib
        import ctypes as C
        for p in self._opened:
            _lib.lib().ecgb_ipc_close(C.c_void_p(p), self.tr.device)
        self._opened = []


def train_shard_persistent(shard, num_merges, exchange=None, device=None, table_log2=0, max_ctas=0, timeout_s=0.0):
    """train_shard with the persistent kernel.  Returns (pairs, counts, ntied, trainer)."""
    n = shard.numel() if isinstance(shard, torch.Tensor) else len(shard)
    st = ShardedTrainer(n, num_merges, exchange=exchange, device=device, table_log2=table_log2)
    try:
        pairs, counts, ntied = st.train(shard, num_merges, max_ctas=max_ctas, timeout_s=timeout_s)
        st.ex.barrier()  # nobody unmaps an area a slower peer may still be writing to
    finally:
        st.close()
    return pairs, counts, ntied, st.tr


# Below this many symbols (whole corpus) a sharded run loses to one GPU: the tail of a small corpus is a chain of
# latencies (argmax, barrier, exchange), not bandwidth.  Measured on B200s: 6e7 symbols -> 2 GPUs are 0.73x of one,
# 1.2e9 symbols -> 1.65x; the switch-over is placed between the two.
SHARD_MIN_SYMBOLS = 250_000_000


def train_corpus_auto(shard, num_merges, exchange=None, device=None, table_log2=0):
    """byte_pair_encoding over a corpus that already lies in contiguous shards on the ranks: sharded
    (ShardedTrainer) when the corpus is large enough to be bandwidth-bound, otherwise gathered onto rank 0,
    trained there by the single-device loop and the merge list broadcast.  -> (pairs, counts, ntied), identical
    on every rank."""
    import torch.distributed as dist
    ex = exchange or TorchExchange()
    dev = shard.device
    n = torch.tensor([shard.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(ex.world)]
    dist.all_gather(sizes, n, group=ex.group)
    sizes = [int(s) for s in sizes]
    total = sum(sizes)
    if ex.world == 1 or total >= SHARD_MIN_SYMBOLS:
        st = ShardedTrainer(shard.numel(), num_merges, exchange=ex, device=device, table_log2=table_log2)
        try:
            res = st.train(shard, num_merges)
            ex.barrier()
        finally:
            st.close()
        return res
    mx = max(sizes)
    pad = torch.zeros(mx, dtype=torch.uint8, device=dev)
    pad[: shard.numel()] = shard
    parts = [torch.zeros(mx, dtype=torch.uint8, device=dev) for _ in range(ex.world)] if ex.rank == 0 else None
    dist.gather(pad, parts, dst=0, group=ex.group)
    m = int(num_merges)
    res = torch.zeros((m, 4), dtype=torch.int64, device=dev)  # left, right, count, ntied
    done = torch.zeros(1, dtype=torch.int64, device=dev)
    if ex.rank == 0:
        corpus = torch.cat([p[:s] for p, s in zip(parts, sizes)])
        tr = Trainer(max(corpus.numel(), 1), m, device=device, table_log2=table_log2)
        tr.load(corpus)
        pairs, counts, ntied = tr.run(m)
        d = len(pairs)
        done[0] = d
        res[:d, :2] = torch.from_numpy(pairs.astype(np.int64)).to(dev)
        res[:d, 2] = torch.from_numpy(counts.astype(np.int64)).to(dev)
        res[:d, 3] = torch.from_numpy(ntied.astype(np.int64)).to(dev)
    dist.broadcast(done, 0, group=ex.group)
    dist.broadcast(res, 0, group=ex.group)
    d = int(done)
    r = res[:d].cpu().numpy()
    return r[:, :2].astype(np.uint32), r[:, 2].astype(np.uint64), r[:, 3].astype(np.uint32)


def train_shards_persistent_local(shards, num_merges, device=None, table_log2=0, max_ctas=0, timeout_s=10.0):
    """The persistent sharded loop with every 'rank' in this process on ONE device, as ONE cooperative launch
    (blockIdx.y = rank, so all ranks are co-resident by construction), exchanging through plain device
    pointers.  Exercises the real protocol -- tagged units, inboxes, shard records -- on a single GPU."""
    import ctypes as C
    from . import _lib
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
    with torch.cuda.device(dev):
        areas = [t.peer_area(world)[0] for t in trs]
        for r, t in enumerate(trs):
            t.dist_begin(r, world, all_bnd[r * bbytes:(r + 1) * bbytes])
        torch.cuda.synchronize(dev)
        for r, t in enumerate(trs):
            t.dist_count(all_bnd, all_lst[r * lbytes:(r + 1) * lbytes])
        torch.cuda.synchronize(dev)
        for t in trs:
            t.dist_apply(all_lst, world)
        torch.cuda.synchronize(dev)
        hs = (C.c_void_p * world)(*[t._h for t in trs])
        ar = (C.c_void_p * world)(*[C.c_void_p(a) for a in areas])
        _lib.check(_lib.lib().ecgb_trainer_dist_run_local(hs, world, ar, C.c_void_p(all_bnd.data_ptr()), int(num_merges),
                                                          int(max_ctas), float(timeout_s),
                                                          C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        res = [t.results(num_merges) for t in trs]
    return res, trs


def _run_steps(one_step, num_merges, dev, use_graph, warm=3):
    """`warm` eager steps (lazy initialisation: NCCL channels, occupancy queries), then -- if allowed --
    one captured step replayed for the rest."""
    eager = min(num_merges, warm)
    for _ in range(eager):
        one_step()
    left = num_merges - eager
    if left <= 0:
        return
    graph = None
    if use_graph:
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):  # capturing does not execute the step
            one_step()
    if graph is not None:
        for _ in range(left):
            graph.replay()
    else:
        for _ in range(left):
            one_step()


def train_shards_local(shards, num_merges, device=None, table_log2=0, use_graph=None):
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
        S = Trainer.STEP_DEVICE
        nb = [nbnd[r * bbytes:(r + 1) * bbytes] for r in range(world)]
        nl = [nlst[r * lbytes:(r + 1) * lbytes] for r in range(world)]

        def one_step():
            for r, t in enumerate(trs):
                t.dist_commit(S, all_lst, nb[r])
            all_bnd.copy_(nbnd)
            for r, t in enumerate(trs):
                t.dist_merge(S, all_bnd, nl[r])
            all_lst.copy_(nlst)
            for t in trs:
                t.dist_advance(dev)

        _run_steps(one_step, num_merges, dev, _graph_default(True) if use_graph is None else use_graph)
        res = [t.results(num_merges) for t in trs]
    del bnd, lst
    return res, trs


def split_contiguous(n, world):
    """[lo, hi) of each rank's piece of an n-byte corpus."""
    return [(n * r // world, n * (r + 1) // world) for r in range(world)]
