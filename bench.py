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
FIXTURE_10K = os.path.join(ROOT, "tests", "golden", "ptbxl_1000_m10000.npz")
WORKLOAD = "encode-only: 100k synthetic PTB-XL-shaped records (12x5000 fp32) per GPU, fixed 5000-merge table"
METRIC = "ECG records tokenized/sec"
UNIT = "records/s"


def bench_config(args, n_merges):
    """The `config` object of the JSON line -- the same keys and values in both arms (the reference arm times a
    bounded sample of this workload and says so in cpu_baseline.sample)."""
    return {"workload": WORKLOAD, "records_per_gpu": int(args.records), "leads": C_LEADS, "samples_per_lead": L_SAMPLES,
            "input_dtype": "fp32", "n_merges": int(n_merges), "table": "tests/golden/ptbxl_1000_m5000.npz (config-1 corpus)",
            "parallelism": "records sharded x%d, no data-path collective" % int(args.gpus),
            "l2": "inputs (%.1f GB/GPU) exceed L2; no flush" % (args.records * REC_LEN * 4 / 1e9)}


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
    # the reference's own NumPy front half as it executes it (np.vectorize, ''.join), one core
    from oracle import py_restatement as P
    t0 = time.perf_counter()
    n_np = 0
    while time.perf_counter() - t0 < 2.0:
        P.normalize_all_as_written(cal[n_np % len(cal)].astype(np.float64), pct["percentile_1"], pct["percentile_99"])
        n_np += 1
    numpy_rate = n_np / (time.perf_counter() - t0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args, len(pairs)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d records/step x %d steps, trie rebuilt per record as rust_bpe.encode_text does" % (per_step, args.steps),
                         "numpy_as_written_records_per_s_per_core": numpy_rate},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- CUDA arm
def bind_to_gpu_numa(dev):
    """CPU affinity of this rank -> the cores next to its GPU, so that the pinned buffers allocated afterwards
    (first touch) and the copy-issuing thread sit on the GPU's NUMA node.  Best effort."""
    try:
        import torch
        p = torch.cuda.get_device_properties(dev)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        path = "/sys/bus/pci/devices/%s/local_cpulist" % bdf
        cpus = set()
        for part in open(path).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "%d cpus of %s" % (len(cpus), bdf)
    except Exception as e:  # noqa: BLE001
        return "not bound (%s)" % type(e).__name__
    return "not bound"


def progress(msg):
    """stage marker on stderr (rank 0): the JSON line on stdout stays alone"""
    if int(os.environ.get("RANK", "0")) == 0:
        sys.stderr.write("[bench %7.1fs] %s\n" % (time.perf_counter() - _T0, msg))
        sys.stderr.flush()


_T0 = time.perf_counter()


def timed(fn, reps, dev, torch):
    """mean milliseconds of fn() over reps calls, CUDA events on the current stream"""
    e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    e[0].record()
    for i in range(reps):
        fn()
        e[i + 1].record()
    torch.cuda.synchronize(dev)
    return [e[i].elapsed_time(e[i + 1]) for i in range(reps)]


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from ecgbyte import synth
    from ecgbyte.api import EncodePipeline, EncodePipelineCSR, Quantizer, Vocab

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(dev)
    pairs, pct = load_table()
    n_rec = args.records
    stride = args.out_stride

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

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

    def allmax(vals):
        if world == 1:
            return [float(a) for a in vals]
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(a) for a in t]

    def allsum(vals):
        if world == 1:
            return [float(a) for a in vals]
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return [float(a) for a in t]

    progress("batch generated; timing the fused encode kernel")
    # ---- headline: the fused quantise+encode kernel over the resident batch ----
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

    progress("encode timed; quantize_kernel alone")
    # ---- K1 alone: ecgb_quantize over the same batch (C*L*(e+1) bytes per record) ----
    sym_out = torch.empty((n_rec, REC_LEN), dtype=torch.uint8, device=dev)
    q.quantize(x, out=sym_out)
    barrier()
    k1_ms = float(np.mean(timed(lambda: q.quantize(x, out=sym_out), 5, dev, torch)))
    k1_bytes = n_rec * REC_LEN * (4 + 1)
    # every record of the timed batch, on the device: the tokens of the fused kernel decode (decode_text, tu.py:75-77)
    # back to exactly the symbols quantize_kernel produces -- K1 and K1-inside-K2 agree and no token is lost or misplaced
    roundtrip_ok = True
    for a in range(0, n_rec, 25000):
        b = min(n_rec, a + 25000)
        dec, dec_len = v.decode_symbols(tokens[a:b], lens[a:b], REC_LEN)
        roundtrip_ok = roundtrip_ok and bool((dec_len == REC_LEN).all()) and bool(torch.equal(dec, sym_out[a:b]))
        del dec, dec_len
    if not roundtrip_ok:
        raise SystemExit("bench.py: PARITY FAILURE (decode(encode(record)) != quantize(record) for some record) -- numbers withheld")
    # spot check of K1 on the timed batch
    from oracle import oracle as O
    k1_idx = [0, n_rec // 2, n_rec - 1]
    k1_ok = all(np.array_equal(sym_out[r].cpu().numpy(),
                               O.quantize(x[r].cpu().numpy(), pct["percentile_1"], pct["percentile_99"]).reshape(-1)) for r in k1_idx)
    del sym_out
    if not k1_ok:
        raise SystemExit("bench.py: PARITY FAILURE (quantize_kernel differs from the oracle) -- numbers withheld")

    progress("end-to-end pipelines (host buffers)")
    # ---- e2e: host buffers through the public host API (H2D + kernels + D2H every step) ----
    n_e2e = min(args.e2e_records, n_rec)
    xh = torch.empty((n_e2e, C_LEADS, L_SAMPLES), dtype=torch.float32).pin_memory()
    xh.copy_(x[:n_e2e])
    tok16_h = torch.empty((n_e2e * stride,), dtype=torch.uint16).pin_memory()
    len_h = torch.empty((n_e2e,), dtype=torch.int32).pin_memory()
    off_h = torch.empty((n_e2e,), dtype=torch.int64).pin_memory()
    pipe = EncodePipelineCSR(v, q, REC_LEN, stride, chunk=args.e2e_chunk, depth=args.e2e_depth)
    e2e_steps = max(3, min(args.steps, 5))

    def time_pipe(p, xin, *outs):
        for _ in range(2):
            p.run(xin, *outs)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n_l = 0
        for _ in range(e2e_steps):
            n_l += p.run(xin, *outs)   # consecutive steps pipeline into each other; every step copies in and out
        e1.record()
        barrier()
        return e0.elapsed_time(e1), n_l

    e2e_ms, launches_e2e = time_pipe(pipe, xh, tok16_h, len_h, off_h)
    assert np.array_equal(len_h.numpy(), lens_h[:n_e2e].astype(np.int32)), "e2e lengths differ from device-resident run"
    e2e_tokens = int(len_h.numpy().astype(np.int64).sum())
    e2e_off = off_h.numpy().copy()
    e2e_tok_copy = [tok16_h[int(e2e_off[k]):int(e2e_off[k]) + int(lens_h[k])].numpy().copy() for k in range(min(8, n_e2e))]

    # the same pipeline fed with int16 records (1 uV/LSB, PTB-XL's native storage type): half the PCIe bytes per record
    e2e_i16_ms = None
    if not args.no_i16:
        q16 = Quantizer(pct, dtype=torch.int16, device=dev)
        x16 = torch.clamp(torch.round(x[:n_e2e] * 1000.0), -32768, 32767).to(torch.int16)
        xh16 = torch.empty((n_e2e, C_LEADS, L_SAMPLES), dtype=torch.int16).pin_memory()
        xh16.copy_(x16)
        pipe16 = EncodePipelineCSR(v, q16, REC_LEN, stride, chunk=args.e2e_chunk, depth=args.e2e_depth)
        e2e_i16_ms, _ = time_pipe(pipe16, xh16, tok16_h, len_h, off_h)
        del x16, pipe16

    # padded int32 rows (round-1 output format), for comparison
    tok32_h = torch.empty((n_e2e, stride), dtype=torch.int32).pin_memory()
    pipe32 = EncodePipeline(v, q, REC_LEN, stride, chunk=args.e2e_chunk, depth=args.e2e_depth)
    e2e32_ms, _ = time_pipe(pipe32, xh, tok32_h, len_h)
    del pipe32

    progress("bare copy ceiling")
    # ---- bare pinned-copy ceiling of this box at `world` concurrent ranks: the same bytes, no kernels ----
    d_buf = torch.empty((n_e2e, C_LEADS, L_SAMPLES), dtype=torch.float32, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def copies(h2d, d2h):
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            if h2d:
                with torch.cuda.stream(s_in):
                    d_buf.copy_(xh, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s_out):
                    tok32_h.copy_(tokens[:n_e2e], non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()
        return (time.perf_counter() - t0) / 3

    copies(True, True)
    t_h2d = copies(True, False)
    t_d2h = copies(False, True)
    t_both = copies(True, True)
    h2d_gbs = n_e2e * REC_LEN * 4 / t_h2d / 1e9
    d2h_gbs = n_e2e * stride * 4 / t_d2h / 1e9
    del d_buf, tok32_h

    # ---- max / sum over ranks ----
    total_ms, e2e_ms, e2e32_ms, k1_ms_max, t_h2d_max = allmax([total_ms, e2e_ms, e2e32_ms, k1_ms, t_h2d])
    if e2e_i16_ms is not None:
        e2e_i16_ms = allmax([e2e_i16_ms])[0]
    total_tokens_all, e2e_tokens_all, h2d_sum, d2h_sum = allsum([total_tokens, e2e_tokens, h2d_gbs, d2h_gbs])

    progress("config 3")
    # ---- config 3: the 10,000-merge table (does not fit in shared memory next to the walkers' rings) ----
    cfg3 = None
    if not args.no_config3:
        cfg3 = bench_config3(args, dev, rank, world, x, q, pct, peak, barrier, allmax)

    del x, tokens
    torch.cuda.empty_cache()

    progress("training")
    # ---- BPE training: config 1 and config 4 ----
    train = None
    if not args.no_train:
        train = bench_train(args, dev, pct, peak, rank, world)

    if rank != 0:
        return

    value = world * n_rec * args.steps / (total_ms * 1e-3)
    e2e_value = world * n_e2e * e2e_steps / (e2e_ms * 1e-3)

    progress("parity gate")
    # ---- parity gate: a sample of the timed batch against the CPU oracle ----
    x = synth.corpus_cuda(2024, n_rec, L_SAMPLES, torch.float32, dev, start=rank * n_rec)
    tokens = torch.empty((n_rec, stride), dtype=torch.int32, device=dev)
    v.encode_batch(q, x, out_stride=stride, tokens=tokens, lens=lens)
    rng = np.random.default_rng(0)
    idx = np.sort(np.concatenate([np.arange(min(8, n_rec)), rng.choice(n_rec, size=min(args.check, n_rec), replace=False)]))
    idx = np.unique(idx)
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
    for k in range(min(8, n_e2e)):  # the host-buffer path (compact 2-byte output) returns the same tokens
        if not np.array_equal(e2e_tok_copy[k].astype(np.int32), g_tok[k, : lens_h[k]]):
            bad.append("e2e tokens of record %d" % k)
    if bad:
        raise SystemExit("bench.py: PARITY FAILURE against the oracle (%s) -- numbers withheld" % ", ".join(bad[:5]))

    # ---- roofline of the (single) kernel of a step ----
    alg_bytes = n_rec * (REC_LEN * 4 + 4) + 4 * total_tokens          # C*L*e + 4T + 4 per record
    k_ms = float(np.mean(kernel_ms))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "encode_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj["dram_bytes_per_record"] * n_rec
            traffic_src = "profiles/encode_traffic.json (ncu --set full, commit %s)" % tj.get("commit", "?")
        except Exception:  # noqa: BLE001
            traffic = None

    progress("CPU baseline")
    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import py_restatement as P
        cores = os.cpu_count() or 1
        # calibrate on a few records, then time ~12 s of CPU work on records of the same batch
        cal_rate, _, _ = cpu_encode_rate(x[: 4 * cores].cpu().numpy(), pct, pairs, cores, faithful=True)
        n_cpu = int(min(max(cal_rate * 12.0, 4 * cores), n_rec, 65536))
        xs_cpu = x[:n_cpu].cpu().numpy()
        rate, dt, _ = cpu_encode_rate(xs_cpu, pct, pairs, cores, faithful=True)
        rate_am, dt_am, _ = cpu_encode_rate(xs_cpu[: max(n_cpu // 4, 4 * cores)], pct, pairs, cores, faithful=False)
        # the reference's NumPy front half exactly as it executes it (np.vectorize + ''.join), one core, ~3 s
        t0 = time.perf_counter()
        n_np = 0
        while time.perf_counter() - t0 < 3.0:
            P.normalize_all_as_written(xs_cpu[n_np % n_cpu].astype(np.float64), pct["percentile_1"], pct["percentile_99"])
            n_np += 1
        np_rate = n_np / (time.perf_counter() - t0)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d records of the same batch in %.1f s; C port of normalize_all + rust_bpe.encode_text "
                         "(trie rebuilt per record as lib.rs:153-161 does); trie built once: %.0f records/s"
                         % (n_cpu, dt, rate_am),
               "numpy_as_written": {"value": np_rate, "unit": "records/s per core", "records": n_np,
                                    "what": "normalize_all + ''.join exactly as tokenizer_utils.py:14-19,59 execute them "
                                            "(np.vectorize lambda per sample), float64 records, BEFORE encode_text is called; "
                                            "x cores if fanned out like tokenizer_utils.py:89-91: %.0f records/s" % (np_rate * cores)}}

    rec_in_bytes, rec_out_bytes = REC_LEN * 4, 2.0 * e2e_tokens_all / (world * n_e2e) + 4
    ceiling = h2d_sum * 1e9 / rec_in_bytes  # records/s if nothing but the input copy existed (the output copy runs the other way)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, len(pairs)),
        "notes": {"arithmetic": "fp32 threshold classification, bit-identical to the reference's float64 expression; "
                                "pair-table trie (4-byte entries, two symbols per probe), int32 tokens",
                  "tokens_per_record": total_tokens_all / (world * n_rec), "out_stride": stride, "host_numa": numa},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(world * n_e2e * rec_in_bytes),
                "d2h_bytes_per_step": int(2 * e2e_tokens_all + 12 * world * n_e2e), "records_per_step": world * n_e2e, "steps": e2e_steps,
                "api": "ecgbyte.api.EncodePipelineCSR.run (pinned host records in; 2-byte token ids stored by the kernel straight "
                       "into pinned host memory, + int64 offsets and int32 lengths; consecutive steps pipeline into each other)",
                "copy_ceiling": {"h2d_gbs_all_ranks": h2d_sum, "d2h_gbs_all_ranks": d2h_sum,
                                 "h2d_plus_d2h_concurrent_s": t_both, "records_per_s": ceiling,
                                 "what": "bare pinned cudaMemcpyAsync of the same buffers on %d rank(s) at once, no kernels" % world},
                "frac_of_copy_ceiling": e2e_value / ceiling,
                "padded_int32_rows_records_per_s": world * n_e2e * e2e_steps / (e2e32_ms * 1e-3)},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "kernel": "ecgb::encode2_kernel<F32> (pair-table walker; the batch fills the chip)" if n_rec >= 512 * torch.cuda.get_device_properties(dev).multi_processor_count else "ecgb::encode_kernel<F32>",
                     "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src},
        "quantize": {"kernel": "ecgb::quantize_kernel<F32>", "ms": k1_ms_max, "records_per_s": world * n_rec / (k1_ms_max * 1e-3),
                     "algorithmic_bytes_per_launch": k1_bytes, "achieved_gbs": k1_bytes / (k1_ms * 1e-3) / 1e9,
                     "frac": k1_bytes / (k1_ms * 1e-3) / 1e9 / peak, "parity": "3 records == oracle"},
        "cpu_baseline": cpu,
        "config3": cfg3,
        "train": train,
        "clocks": clocks,
        "parity": {"records_checked": int(len(idx)), "ok": True,
                   "what": "tokens of %d records of the timed batch == oracle; all %d records: decode(tokens) == quantize_kernel symbols "
                           "(device-side round trip); e2e tokens == resident tokens; K1 symbols, merge lists == oracle" % (len(idx), n_rec)},
    }
    if e2e_i16_ms is not None:
        i16_value = world * n_e2e * e2e_steps / (e2e_i16_ms * 1e-3)
        line["e2e_int16"] = {"value": i16_value, "unit": UNIT, "h2d_bytes_per_step": int(world * n_e2e * REC_LEN * 2),
                             "d2h_bytes_per_step": int(2 * e2e_tokens_all + 12 * world * n_e2e),
                             "what": "the same pipeline fed int16 records (1 uV/LSB, PTB-XL's storage type)",
                             "frac_of_copy_ceiling": i16_value / (h2d_sum * 1e9 / (REC_LEN * 2))}
    print(json.dumps(line), flush=True)


def bench_config3(args, dev, rank, world, x, q, pct, peak, barrier, allmax):
    """BASELINE config 3: encode against a 10,000-merge table, a 1M-record corpus sharded over the GPUs.  The table is
    trained here by the GPU trainer on the config-1 corpus and must equal the oracle's (fixture); each rank's share of
    the 1M records is covered by repeated passes over its resident 100k-record batch."""
    import torch
    from ecgbyte import synth
    from ecgbyte.api import Quantizer, Trainer, Vocab
    from oracle import oracle as O
    f = np.load(FIXTURE_10K)
    want = f["pairs"].astype(np.uint32)
    xc = torch.from_numpy(synth.corpus(0, 1000, L_SAMPLES, np.float32)).to(dev)
    symc = q.quantize(xc).reshape(-1)
    tr = Trainer(symc.numel(), len(want), device=dev)
    tr.load(symc)
    t0 = time.perf_counter()
    pairs10, counts10, _ = tr.run(len(want))
    t_train = time.perf_counter() - t0
    del tr, xc, symc
    if not (np.array_equal(pairs10, want) and np.array_equal(counts10, f["counts"])):
        raise SystemExit("bench.py: PARITY FAILURE (10,000-merge table differs from the oracle fixture)")
    v10 = Vocab.from_pairs(pairs10, device=dev)
    n_rec = x.shape[0]
    stride = args.out_stride
    tokens = torch.empty((n_rec, stride), dtype=torch.int32, device=dev)
    lens = torch.empty((n_rec,), dtype=torch.int32, device=dev)
    share = (1000000 + world - 1) // world
    passes = max(1, (share + n_rec - 1) // n_rec)
    for _ in range(3):
        v10.encode_batch(q, x, out_stride=stride, tokens=tokens, lens=lens)
    barrier()
    ms = timed(lambda: v10.encode_batch(q, x, out_stride=stride, tokens=tokens, lens=lens), passes, dev, torch)
    barrier()
    tot_ms = allmax([float(np.sum(ms))])[0]
    T = int(lens.sum().item())
    out = None
    if rank == 0:
        idx = np.array([0, n_rec // 3, n_rec - 1])
        xs = x[torch.from_numpy(idx).to(dev)].cpu().numpy()
        sym = O.quantize(xs, pct["percentile_1"], pct["percentile_99"]).reshape(len(idx), -1)
        seq, off = O.expand(pairs10)
        trie = O.Trie(flat=(seq, off, np.arange(256, 256 + len(pairs10), dtype=np.uint32)))
        w_tok, w_len = trie.encode_batch(sym, stride)
        g_tok = tokens[torch.from_numpy(idx).to(dev)].cpu().numpy()
        g_len = lens.cpu().numpy()[idx]
        ok = np.array_equal(w_len.astype(np.int64), g_len.astype(np.int64)) and all(
            np.array_equal(g_tok[k, : w_len[k]], w_tok[k, : w_len[k]].astype(np.int32)) for k in range(len(idx)))
        if not ok:
            raise SystemExit("bench.py: PARITY FAILURE (config 3 tokens differ from the oracle)")
        alg = n_rec * (REC_LEN * 4 + 4) + 4 * T
        k_ms = float(np.mean(ms))
        out = {"workload": "encode-only: 1M synthetic records (12x5000 fp32) sharded over %d GPU(s), 10,000-merge table" % world,
               "value": world * passes * n_rec / (tot_ms * 1e-3), "unit": UNIT, "records": world * passes * n_rec,
               "passes_over_resident_batch": passes, "records_per_gpu_resident": n_rec, "ms_per_pass": k_ms,
               "tokens_per_record": T / n_rec, "table": dict(v10.info()),
               "roofline": {"bound": "hbm", "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": alg / (k_ms * 1e-3) / 1e9 / peak},
               "table_training": {"merges": int(len(pairs10)), "seconds": t_train, "parity": "merge list == oracle fixture (10000 merges)"},
               "parity": "3 records == oracle"}
    del tokens, lens, v10
    return out


def bench_train(args, dev, pct, peak, rank, world):
    """BPE-train merges/s.  Config 1 (1,000 records = 6e7 symbols, 5,000 merges): one GPU runs train_loop_kernel; N GPUs
    run dist_loop_kernel on N contiguous shards (device-initiated exchange over NVLink), and the `auto` policy
    (corpus fits one GPU -> do not shard) is reported beside it.  Config 4 (250k records = 1.5e10 symbols, 20,000
    merges) at the same N.  Merge lists are checked against the oracle fixture (config 1) and, for config 4, on a
    CPU-holdable sub-corpus; the config-4 merge list's CRC is printed so that runs at different N can be compared."""
    import torch
    import torch.distributed as dist
    from ecgbyte import synth
    from ecgbyte.api import Quantizer, Trainer
    from ecgbyte.dist_train import ShardedTrainer, split_contiguous
    from oracle import oracle as O
    f = np.load(FIXTURE)
    q = Quantizer(pct, dtype=torch.float32, device=dev)
    out = {}

    def crc(p):
        p = np.asarray(p, np.uint64).reshape(-1)
        return int(np.bitwise_xor.reduce(p * (np.arange(1, len(p) + 1, dtype=np.uint64) * np.uint64(2654435761))))

    def shard_symbols(seed, n_total, lo, hi):
        s = torch.empty((hi - lo) * REC_LEN, dtype=torch.uint8, device=dev)
        for a in range(lo, hi, 2048):
            b = min(hi, a + 2048)
            s[(a - lo) * REC_LEN:(b - lo) * REC_LEN] = q.quantize(
                synth.corpus_cuda_range(seed, n_total, a, b, L_SAMPLES, torch.float32, dev)).reshape(-1)
        return s

    def run(shard, m, tlog, reps, st=None):
        """-> (best seconds (max over ranks), pairs, counts, per-step stream lengths summed over ranks)"""
        best = None
        for _ in range(reps):
            if world == 1:
                tr = st if st is not None else Trainer(shard.numel(), m, device=dev, table_log2=tlog)
                tr.load(shard)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                pairs, counts, ntied = tr.run(m)
                dt = time.perf_counter() - t0
                lens = tr.lengths(len(pairs)).astype(np.float64)
                st = tr
            else:
                if st is None:
                    st = ShardedTrainer(shard.numel(), m, table_log2=tlog)
                dist.barrier()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                pairs, counts, ntied = st.train(shard, m)
                d = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
                dist.all_reduce(d, op=dist.ReduceOp.MAX)
                dt = float(d)
                l = torch.from_numpy(st.tr.lengths(len(pairs)).astype(np.float64)).to(dev)
                dist.all_reduce(l)
                lens = l.cpu().numpy()
            best = dt if best is None else min(best, dt)
        if world > 1:
            dist.barrier()
            st.close()
        return best, pairs, counts, lens

    progress("training: config 1")
    # ---- config 1 ----
    xc = synth.corpus(0, 1000, L_SAMPLES, np.float32)
    sym = q.quantize(torch.from_numpy(xc).to(dev)).reshape(-1)
    n1 = sym.numel()
    lo, hi = split_contiguous(n1, world)[rank]
    m1 = 5000
    t1, pairs, counts, lens1 = run(sym[lo:hi].contiguous(), m1, 0, 3)
    if not (np.array_equal(pairs, f["pairs"].astype(np.uint32)) and np.array_equal(counts, f["counts"])):
        raise SystemExit("bench.py: PARITY FAILURE (merge list differs from the oracle fixture)")
    alg1 = float(np.sum(2.0 * (lens1[:-1] + lens1[1:])))  # 2*(n_t + n_{t+1}) bytes per step
    out = {"metric": "BPE-train merges/sec", "value": m1 / t1, "unit": "merges/s", "seconds": t1,
           "config": "config 1: 1,000 records = %d symbols, %d merges" % (n1, m1), "gpus": world,
           "kernel": "train_loop_kernel (1 GPU, persistent)" if world == 1 else
                     "dist_loop_kernel x%d (persistent, shard records + histogram patches written into the peers' memory over NVLink)" % world,
           "final_tokens": int(lens1[-1]), "algorithmic_bytes": alg1, "achieved_gbs": alg1 / t1 / 1e9,
           "frac_of_hbm_peak": alg1 / t1 / 1e9 / (peak * world), "gpu_launches": 2 if world == 1 else 6,
           "parity": "merge list == oracle fixture (5000 merges)"}
    if world > 1:
        # the public entry point for a corpus that lies in shards on the ranks applies the auto policy: below
        # SHARD_MIN_SYMBOLS the shards are gathered on rank 0, trained there by the single-device loop (the tail of a
        # small corpus is a chain of latencies, and one GPU has the shortest chain) and the merges broadcast
        from ecgbyte.dist_train import SHARD_MIN_SYMBOLS, train_corpus_auto
        t_auto = None
        for _ in range(3):
            dist.barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            pa, ca, _ = train_corpus_auto(sym[lo:hi].contiguous(), m1)
            d = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            dist.all_reduce(d, op=dist.ReduceOp.MAX)
            t_auto = float(d) if t_auto is None else min(t_auto, float(d))
        if not (np.array_equal(pa, pairs) and np.array_equal(ca, counts)):
            raise SystemExit("bench.py: PARITY FAILURE (auto-policy merge list differs)")
        out["sharded_forced"] = {k: out[k] for k in ("value", "unit", "seconds", "kernel", "achieved_gbs", "frac_of_hbm_peak", "gpu_launches")}
        out.update({"value": m1 / t_auto, "seconds": t_auto, "achieved_gbs": alg1 / t_auto / 1e9,
                    "frac_of_hbm_peak": alg1 / t_auto / 1e9 / peak, "gpu_launches": 2,
                    "kernel": "ecgbyte.dist_train.train_corpus_auto: %d symbols < SHARD_MIN_SYMBOLS (%d) -> shards gathered over NCCL, "
                              "train_loop_kernel on rank 0, merge list broadcast (gather and broadcast inside the timed region)"
                              % (n1, SHARD_MIN_SYMBOLS)})
    del sym

    # ---- config 4 ----
    if not args.no_config4:
        progress("training: config-4 sub-corpus parity")
        # (a) parity on a CPU-holdable sub-corpus through the same code path
        n_sub, m_sub = 200, 300
        lo, hi = split_contiguous(n_sub, world)[rank]
        sub = shard_symbols(0, n_sub, lo, hi)
        _, p_sub, c_sub, _ = run(sub, m_sub, 0, 1)
        if rank == 0:
            full = shard_symbols(0, n_sub, 0, n_sub).cpu().numpy()
            _, o_pairs, o_counts, _ = O.train_pairs(full, m_sub, fast=True)
            if not (np.array_equal(p_sub, o_pairs) and np.array_equal(c_sub, o_counts)):
                raise SystemExit("bench.py: PARITY FAILURE (config-4 sub-corpus merge list differs from the oracle)")
        del sub
        progress("training: config 4, generating the corpus")
        # (b) the full corpus
        n4, m4 = args.config4_records, args.config4_merges
        lo, hi = split_contiguous(n4, world)[rank]
        t0 = time.perf_counter()
        shard = shard_symbols(0, n4, lo, hi)
        torch.cuda.synchronize(dev)
        gen = time.perf_counter() - t0
        progress("training: config 4, %d symbols on this rank" % shard.numel())
        t4, p4, c4, lens4 = run(shard, m4, 26, 1)
        progress("training: config 4 done in %.1f s" % t4)
        del shard
        alg4 = float(np.sum(2.0 * (lens4[:-1] + lens4[1:])))
        out["config4"] = {"config": "config 4: %d records = %d symbols, %d merges" % (n4, int(lens4[0]), m4), "gpus": world,
                          "value": len(p4) / t4, "unit": "merges/s", "seconds": t4, "merges_done": int(len(p4)),
                          "final_tokens": int(lens4[-1]), "algorithmic_bytes": alg4, "achieved_gbs": alg4 / t4 / 1e9,
                          "frac_of_hbm_peak": alg4 / t4 / 1e9 / (peak * world), "generate_seconds": gen,
                          "merge_list_crc": crc(p4), "first_merges": np.asarray(p4[:3]).tolist(), "last_merge": np.asarray(p4[-1]).tolist(),
                          "top_count": int(c4[0]),
                          "parity": "sub-corpus (%d records, %d merges) through the same path == oracle; compare merge_list_crc across N"
                                    % (n_sub, m_sub)}
    return out


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
    ap.add_argument("--e2e-depth", type=int, default=4, help="chunks in flight in the end-to-end pipeline")
    ap.add_argument("--check", type=int, default=64)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-i16", action="store_true")
    ap.add_argument("--no-config3", action="store_true")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--config4-records", type=int, default=250000)
    ap.add_argument("--config4-merges", type=int, default=20000)
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
