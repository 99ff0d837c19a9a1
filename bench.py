#!/usr/bin/env python
"""Benchmark of the ECG-Byte tokenizer hot path (BASELINE.json metric:
"ECG samples tokenized/sec"; a sample = one 12-lead 500 Hz 10 s record).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, libecgbyte.so)
  python bench.py --impl reference --gpus N --steps K ...   the reference's CPU path

Workload (BASELINE.json configs[1]): encode-only, 100k synthetic PTB-XL-shaped records
(12 x 5000 fp32) per GPU against a fixed 5,000-merge table trained on the 1,000-record
config-1 corpus (tests/golden/ptbxl_1000_m5000.npz).  A step = one fused
quantise+encode pass over the resident batch (24 GB per GPU, far larger than the 126 MB
L2, so no cache flush is needed between steps).  Records shard across ranks with no
data-path collective (weak scaling: 100k records per GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "ecg-byte_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

C_LEADS, L_SAMPLES = 12, 5000
REC_LEN = C_LEADS * L_SAMPLES
FIXTURE = os.path.join(ROOT, "tests", "golden", "ptbxl_1000_m5000.npz")
WORKLOAD = "encode-only: 100k synthetic PTB-XL-shaped records (12x5000 fp32) per GPU, fixed 5000-merge table"
METRIC = "ECG records tokenized/sec"
UNIT = "records/s"


def load_table():
    f = np.load(FIXTURE)
    pct = {"percentile_1": np.float64(f["pct"][0]), "percentile_99": np.float64(f["pct"][1])}
    return f["pairs"].astype(np.uint32), pct


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arm
def cpu_encode_rate(x_host, pct, pairs, threads, faithful=True):
    """The reference's CPU path restated in C (oracle): per record, quantise
    (tokenizer_utils.py:14-19) then encode_text with the trie REBUILT on every call, as
    rust_bpe does (lib.rs:153-161), records fanned out over `threads` host threads
    (the reference fans out over processes, tokenizer_utils.py:89-91)."""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    L = O.lib()
    seq, off = O.expand(pairs)
    ids = np.arange(256, 256 + len(pairs), dtype=np.uint32)
    n = x_host.shape[0]
    flat = x_host.reshape(n, -1)
    trie = None if faithful else O.Trie(flat=(seq, off, ids))
    p1, p99 = float(pct["percentile_1"]), float(pct["percentile_99"])

    def one(r):
        sym = np.empty(flat.shape[1], np.uint8)
        L.ecgo_quantize(flat[r].ctypes.data, O._DT[flat.dtype], flat.shape[1], p1, p99, 1e-3, sym.ctypes.data)
        out = np.empty(flat.shape[1], np.uint32)
        n_out = C.c_size_t(0)
        if faithful:
            L.ecgo_encode(sym.ctypes.data, sym.size, seq.ctypes.data, off.ctypes.data, ids.ctypes.data, len(pairs),
                          out.ctypes.data, out.size, C.byref(n_out))
        else:
            L.ecgo_trie_encode(trie.h, sym.ctypes.data, sym.size, out.ctypes.data, out.size, C.byref(n_out))
        return n_out.value

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        toks = list(ex.map(one, range(n)))
    dt = time.perf_counter() - t0
    return n / dt, dt, int(sum(toks))


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle port;
    the real crate needs a Rust toolchain that this image does not have)."""
    if rank != 0:
        return
    from ecgbyte import synth
    pairs, pct = load_table()
    cores = os.cpu_count() or 1
    # a step = a bounded sample of the workload: ~3 s of work on all host cores
    cal = synth.corpus(1234, 4 * cores, L_SAMPLES, np.float32)
    cal_rate, _, _ = cpu_encode_rate(cal, pct, pairs, cores)
    per_step = int(min(max(cal_rate * 3.0, 4 * cores), 8192))
    x = np.concatenate([cal] * ((per_step + len(cal) - 1) // len(cal)))[:per_step]
    for _ in range(min(args.warmup, 1)):
        cpu_encode_rate(x[: max(cores, 8)], pct, pairs, cores)
    t_tot, n_tot = 0.0, 0
    for _ in range(args.steps):
        rate, dt, _ = cpu_encode_rate(x, pct, pairs, cores)
        t_tot += dt
        n_tot += per_step
    value = n_tot / t_tot
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "records_per_step": per_step, "input": "fp32 12x5000", "n_merges": len(pairs)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d records/step x %d steps, trie rebuilt per record as rust_bpe.encode_text does" % (per_step, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- CUDA arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from ecgbyte import synth
    from ecgbyte.api import EncodePipeline, Quantizer, Vocab

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pairs, pct = load_table()
    n_rec = args.records
    stride = args.out_stride

    q = Quantizer(pct, dtype=torch.float32, device=dev)
    v = Vocab.from_pairs(pairs, device=dev)
    x = synth.corpus_cuda(2024, n_rec, L_SAMPLES, torch.float32, dev, start=rank * n_rec)
    tokens = torch.empty((n_rec, stride), dtype=torch.int32, device=dev)
    lens = torch.empty((n_rec,), dtype=torch.int32, device=dev)
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        v.encode_batch(q, x, out_stride=stride, tokens=tokens, lens=lens)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        v.encode_batch(q, x, out_stride=stride, tokens=tokens, lens=lens)
        ev[i + 1].record()
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if rank == 0 else None

    lens_h = lens.cpu().numpy().astype(np.int64)
    assert lens_h.max() <= stride, "out_stride %d too small (max tokens %d)" % (stride, lens_h.max())
    total_tokens = int(lens_h.sum())

    # ---- e2e: host buffers through the public host API (H2D + kernel + D2H every step) ----
    n_e2e = min(args.e2e_records, n_rec)
    xh = torch.empty((n_e2e, C_LEADS, L_SAMPLES), dtype=torch.float32).pin_memory()
    xh.copy_(x[:n_e2e])
    tok_h = torch.empty((n_e2e, stride), dtype=torch.int32).pin_memory()
    len_h = torch.empty((n_e2e,), dtype=torch.int32).pin_memory()
    pipe = EncodePipeline(v, q, REC_LEN, stride, chunk=args.e2e_chunk)
    for _ in range(2):
        pipe.run(xh, tok_h, len_h)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, min(args.steps, 5))
    e0.record()
    launches_e2e = 0
    for _ in range(e2e_steps):
        launches_e2e += pipe.run(xh, tok_h, len_h)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    assert np.array_equal(len_h.numpy(), lens_h[:n_e2e].astype(np.int32)), "e2e lengths differ from device-resident run"

    # informational: the same pipeline fed with int16 records (1 uV/LSB, PTB-XL's native type):
    # half the PCIe bytes per record
    e2e_i16 = None
    if not args.no_i16:
        q16 = Quantizer(pct, dtype=torch.int16, device=dev)
        x16 = torch.clamp(torch.round(x[:n_e2e] * 1000.0), -32768, 32767).to(torch.int16)
        xh16 = torch.empty((n_e2e, C_LEADS, L_SAMPLES), dtype=torch.int16).pin_memory()
        xh16.copy_(x16)
        pipe16 = EncodePipeline(v, q16, REC_LEN, stride, chunk=args.e2e_chunk)
        pipe16.run(xh16, tok_h, len_h)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(e2e_steps):
            pipe16.run(xh16, tok_h, len_h)
        f1.record()
        barrier()
        e2e_i16 = n_e2e * e2e_steps / (f0.elapsed_time(f1) * 1e-3)
        del x16, xh16, pipe16
        # restore the fp32 results in the pinned buffers for the parity gate below
        pipe.run(xh, tok_h, len_h)
        torch.cuda.synchronize(dev)

    # ---- max over ranks ----
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
        tt = torch.tensor([total_tokens], dtype=torch.int64, device=dev)
        dist.all_reduce(tt)
        total_tokens_all = int(tt[0])
    else:
        total_tokens_all = total_tokens
    train_dist = None
    if world > 1 and not args.no_train:
        train_dist = bench_train_sharded(dev, pct, rank, world)
    if rank != 0:
        return

    value = world * n_rec * args.steps / (total_ms * 1e-3)
    e2e_value = world * n_e2e * e2e_steps / (e2e_ms * 1e-3)

    # ---- parity gate: a sample of the timed batch against the CPU oracle ----
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    idx = np.sort(rng.choice(n_rec, size=min(args.check, n_rec), replace=False))
    xs = x[torch.from_numpy(idx).to(dev)].cpu().numpy()
    sym = O.quantize(xs, pct["percentile_1"], pct["percentile_99"]).reshape(len(idx), -1)
    seq, off = O.expand(pairs)
    trie = O.Trie(flat=(seq, off, np.arange(256, 256 + len(pairs), dtype=np.uint32)))
    w_tok, w_len = trie.encode_batch(sym, stride)
    g_tok = tokens[torch.from_numpy(idx).to(dev)].cpu().numpy()
    bad = []
    if not np.array_equal(w_len.astype(np.int64), lens_h[idx]):
        bad.append("token counts")
    for k in range(len(idx)):
        if not np.array_equal(g_tok[k, : w_len[k]], w_tok[k, : w_len[k]].astype(np.int32)):
            bad.append("tokens of record %d" % idx[k])
    d4 = tokens[:4].cpu().numpy()
    for k in range(4):  # the host-buffer path returns the same tokens as the resident path
        if not np.array_equal(tok_h[k, : lens_h[k]].numpy(), d4[k, : lens_h[k]]):
            bad.append("e2e tokens of record %d" % k)
    if bad:
        raise SystemExit("bench.py: PARITY FAILURE against the oracle (%s) -- numbers withheld" % ", ".join(bad[:5]))

    # ---- roofline of the (single) kernel of a step ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = n_rec * (REC_LEN * 4 + 4) + 4 * total_tokens          # C*L*e + 4T + 4 per record
    k_ms = float(np.mean(kernel_ms))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "encode_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj["dram_bytes_per_record"] * n_rec
        except Exception:
            traffic = None

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        # calibrate on a few records, then time ~12 s of CPU work on records of the same batch
        cal_rate, _, _ = cpu_encode_rate(x[: 4 * cores].cpu().numpy(), pct, pairs, cores, faithful=True)
        n_cpu = int(min(max(cal_rate * 12.0, 4 * cores), n_rec, 65536))
        xs_cpu = x[:n_cpu].cpu().numpy()
        rate, dt, _ = cpu_encode_rate(xs_cpu, pct, pairs, cores, faithful=True)
        rate_am, dt_am, _ = cpu_encode_rate(xs_cpu[: max(n_cpu // 4, 4 * cores)], pct, pairs, cores, faithful=False)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d records of the same batch in %.1f s; C port of normalize_all + rust_bpe.encode_text "
                         "(trie rebuilt per record as lib.rs:153-161 does); trie built once: %.0f records/s"
                         % (n_cpu, dt, rate_am)}

    # ---- secondary metric: BPE-train merges/s (BASELINE.json config 1 scale), single GPU ----
    train = None
    if world == 1 and not args.no_train:
        train = bench_train(dev, pct, peak)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "records_per_gpu": n_rec, "leads": C_LEADS, "samples_per_lead": L_SAMPLES,
                   "input_dtype": "fp32", "n_merges": int(len(pairs)), "out_stride": stride,
                   "arithmetic": "fp32 threshold classification, bit-identical to the reference's float64 expression; "
                                 "u8 symbols, 8-byte trie nodes, int32 tokens",
                   "tokens_per_record": total_tokens_all / (world * n_rec), "parallelism": "records sharded x%d" % world,
                   "l2": "inputs (%.1f GB/GPU) exceed L2; no flush" % (n_rec * REC_LEN * 4 / 1e9)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * n_e2e * REC_LEN * 4,
                "d2h_bytes_per_step": world * n_e2e * (stride * 4 + 4), "records_per_step": world * n_e2e, "steps": e2e_steps,
                "api": "ecgbyte.api.EncodePipeline.run (pinned host in/out)",
                "int16_input_records_per_s_per_gpu": e2e_i16},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "ecgb::encode_kernel<F32>", "kernel_ms": k_ms,
                     "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src},
        "cpu_baseline": cpu,
        "train": train if train is not None else train_dist,
        "clocks": clocks,
        "parity": {"records_checked": int(len(idx)), "ok": True},
    }
    print(json.dumps(line), flush=True)


def bench_train(dev, pct, peak):
    """byte_pair_encoding on the config-1 corpus shape: 1,000 records = 6e7 symbols,
    5,000 merges, whole loop on the device; checked against the oracle's merge list."""
    import torch
    from ecgbyte import synth
    from ecgbyte.api import Quantizer, Trainer
    f = np.load(FIXTURE)
    x = torch.from_numpy(synth.corpus(0, 1000, L_SAMPLES, np.float32)).to(dev)
    q = Quantizer(pct, dtype=torch.float32, device=dev)
    sym = q.quantize(x).reshape(-1)
    m = 5000
    tr = Trainer(sym.numel(), m, device=dev)
    best = None
    for _ in range(3):
        tr.load(sym)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        pairs, counts, ntied = tr.run(m)  # synchronises at the end
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    ok = bool(np.array_equal(pairs, f["pairs"].astype(np.uint32)) and np.array_equal(counts, f["counts"]))
    if not ok:
        raise SystemExit("bench.py: PARITY FAILURE (merge list differs from the oracle fixture)")
    n = tr.lengths(m).astype(np.float64)
    alg = float(np.sum(2.0 * (n[:-1] + n[1:])))  # 2*(n_t + n_{t+1}) bytes per step
    return {"metric": "BPE-train merges/sec", "value": m / best, "unit": "merges/s", "seconds": best,
            "corpus_symbols": int(n[0]), "merges": m, "final_tokens": int(n[-1]),
            "algorithmic_bytes": alg, "achieved_gbs": alg / best / 1e9, "frac_of_hbm_peak": alg / best / 1e9 / peak,
            "gpu_launches": 2, "parity": "merge list == oracle fixture (5000 merges)"}


def bench_train_sharded(dev, pct, rank, world):
    """The same corpus cut into `world` contiguous shards, one per rank; two NCCL all-gathers
    per merge step (ecgbyte/dist_train.py).  Rank 0 checks the merge list against the fixture."""
    import torch
    import torch.distributed as dist
    from ecgbyte import synth
    from ecgbyte.api import Quantizer
    from ecgbyte.dist_train import split_contiguous, train_shard
    f = np.load(FIXTURE)
    x = torch.from_numpy(synth.corpus(0, 1000, L_SAMPLES, np.float32)).to(dev)
    q = Quantizer(pct, dtype=torch.float32, device=dev)
    sym = q.quantize(x).reshape(-1)
    lo, hi = split_contiguous(sym.numel(), world)[rank]
    shard = sym[lo:hi].contiguous()
    m = 5000
    best = None
    for _ in range(2):
        dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        pairs, counts, ntied, tr = train_shard(shard, m, device=dev)
        torch.cuda.synchronize(dev)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        best = float(dt) if best is None else min(best, float(dt))
    ok = bool(np.array_equal(pairs, f["pairs"].astype(np.uint32)) and np.array_equal(counts, f["counts"]))
    if not ok:
        raise SystemExit("bench.py: PARITY FAILURE (sharded merge list differs from the oracle fixture)")
    return {"metric": "BPE-train merges/sec", "value": m / best, "unit": "merges/s", "seconds": best,
            "corpus_symbols": int(sym.numel()), "merges": m, "shards": world,
            "exchange": "2 NCCL all-gathers per merge step (boundary records, histogram delta lists)",
            "parity": "merge list == oracle fixture (5000 merges)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--records", type=int, default=100000, help="records per GPU per step")
    ap.add_argument("--out-stride", type=int, default=8192)
    ap.add_argument("--e2e-records", type=int, default=16384)
    ap.add_argument("--e2e-chunk", type=int, default=2048)
    ap.add_argument("--check", type=int, default=64)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-i16", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
