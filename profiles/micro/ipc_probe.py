"""torchrun --nproc-per-node 2 profiles/micro/ipc_probe.py : CUDA IPC between ranks + NVLink flag latency."""
import ctypes, os, sys
import torch, torch.distributed as dist

here = os.path.dirname(os.path.abspath(__file__))
L = ctypes.CDLL(os.path.join(here, "libipc_probe.so"))
L.probe_err.restype = ctypes.c_char_p
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

def ck(e, what):
    if e:
        raise SystemExit("%s: %s" % (what, L.probe_err(e).decode()))

size = 64 << 20
p = ctypes.c_void_p()
ck(L.probe_alloc(ctypes.c_size_t(size), ctypes.byref(p)), "alloc")
h = (ctypes.c_ubyte * 64)()
ck(L.probe_export(p, h), "export")
handles = [None] * world
dist.all_gather_object(handles, bytes(h))
peers = {}
for r in range(world):
    if r == rank:
        continue
    q = ctypes.c_void_p()
    ck(L.probe_open(handles[r], ctypes.byref(q)), "open rank %d" % r)
    peers[r] = q
dist.barrier()
if world >= 2 and rank < 2:
    us = ctypes.c_double()
    ck(L.probe_pingpong(p, peers[1 - rank], rank, 20000, ctypes.byref(us)), "pingpong")
    print("rank %d: flag round trip %.2f us (one way ~%.2f us)" % (rank, us.value, us.value / 2), flush=True)
dist.barrier()
gbs = ctypes.c_double()
tgt = peers[(rank + 1) % world]
ck(L.probe_push(ctypes.c_void_p(tgt.value + (32 << 20)), ctypes.c_size_t(16 << 20), 20, ctypes.byref(gbs)), "push")
print("rank %d: 16 MiB remote stores to rank %d: %.1f GB/s" % (rank, (rank + 1) % world, gbs.value), flush=True)
dist.barrier()
dist.destroy_process_group()
