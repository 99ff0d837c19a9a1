"""Seeded synthetic PTB-XL / MIMIC-IV-ECG shaped records.

Record k of corpus `seed` is generated from SeedSequence([seed, k]) so any shard of
the corpus can be regenerated independently on any rank.  Shape (12, L), 500 Hz.
The stored layout is the reference's: one (12, seg_len) array per record,
lead-major (preprocess_utils.py:220-223), float64 there; fp32 (headline), fp64 and
int16 (1 uV/LSB) variants here.
"""
import numpy as np

FS = 500.0
N_LEADS = 12
# (amplitude factor, centre in beat phase, sigma in beat phase): P, Q, R, T
_WAVES = ((0.12, 0.33, 0.03), (-0.25, 0.47, 0.01), (1.0, 0.50, 0.012), (0.3, 0.72, 0.05))


def record(seed, k, L=5000, dtype=np.float32):
    rng = np.random.default_rng(np.random.SeedSequence([int(seed), int(k)]))
    hr = rng.uniform(50.0, 110.0)
    phase0 = rng.uniform(0.0, 1.0)
    amp = rng.normal(1.0, 0.4, size=(N_LEADS, 1))
    wander_phi = rng.uniform(0.0, 2 * np.pi, size=(N_LEADS, 1))
    t = np.arange(L, dtype=np.float64) / FS
    beat = np.mod(t * (hr / 60.0) + phase0, 1.0)[None, :]
    x = np.zeros((N_LEADS, L), np.float64)
    for a, c, s in _WAVES:
        x += (a * amp) * np.exp(-0.5 * ((beat - c) / s) ** 2)
    x += 0.05 * np.sin(2 * np.pi * 0.3 * t[None, :] + wander_phi)
    x += rng.normal(0.0, 0.01, size=(N_LEADS, L))
    return cast(x, dtype)


def cast(x_mv, dtype):
    dtype = np.dtype(dtype)
    if dtype == np.int16:
        return np.clip(np.rint(x_mv * 1000.0), -32768, 32767).astype(np.int16)
    return x_mv.astype(dtype)


def corpus(seed, n, L=5000, dtype=np.float32, start=0):
    out = np.empty((n, N_LEADS, L), dtype)
    for i in range(n):
        out[i] = record(seed, start + i, L, dtype)
    return out


def percentiles(records_mv, sample_size=100000, seed=0):
    """1st / 99th percentile of `sample_size` uniformly sampled values, returned as
    the stats dict the reference stores (preprocess_utils.py:197-210)."""
    flat = np.asarray(records_mv).reshape(-1)
    rng = np.random.default_rng(np.random.SeedSequence([int(seed), 0x9e3779b9]))
    idx = rng.integers(0, flat.size, size=min(sample_size, flat.size))
    s = flat[idx].astype(np.float64)
    return {
        "global_min": np.float64(flat.min()),
        "global_max": np.float64(flat.max()),
        "percentile_1": np.percentile(s, 1),
        "percentile_99": np.percentile(s, 99),
        "skipped_instances": 0,
    }


# Fixed stats for the benchmark corpora (computed once from corpus(seed=0, n=1000);
# kept literal so every rank / run quantises identically without a data pass).
BENCH_PERCENTILES = {"percentile_1": np.float64(-0.1610814356803894), "percentile_99": np.float64(0.9105487620830528)}


def corpus_cuda(seed, n, L=5000, dtype=None, device=None, start=0, chunk=2048):
    """Same distribution as corpus(), generated on the device with torch (bulk benchmark
    data: 100k records = 24 GB would take minutes in NumPy).  Record streams are not
    bit-identical to record(); parity checks read the generated tensors back."""
    import torch
    dtype = dtype or torch.float32
    device = torch.device(device if device is not None else "cuda")
    out = torch.empty((n, N_LEADS, L), dtype=dtype, device=device)
    t = torch.arange(L, dtype=torch.float32, device=device) / FS
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        g = torch.Generator(device=device)
        g.manual_seed((int(seed) << 32) + start + c0)
        hr = torch.empty((m, 1, 1), device=device).uniform_(50.0, 110.0, generator=g)
        ph = torch.empty((m, 1, 1), device=device).uniform_(0.0, 1.0, generator=g)
        amp = torch.empty((m, N_LEADS, 1), device=device).normal_(1.0, 0.4, generator=g)
        wphi = torch.empty((m, N_LEADS, 1), device=device).uniform_(0.0, 6.283185307, generator=g)
        beat = torch.remainder(t.view(1, 1, L) * (hr / 60.0) + ph, 1.0)
        x = torch.zeros((m, N_LEADS, L), device=device)
        for a, c, s in _WAVES:
            x += (a * amp) * torch.exp(-0.5 * ((beat - c) / s) ** 2)
        x += 0.05 * torch.sin(6.283185307 * 0.3 * t.view(1, 1, L) + wphi)
        x += torch.empty((m, N_LEADS, L), device=device).normal_(0.0, 0.01, generator=g)
        if dtype == torch.int16:
            out[c0:c0 + m] = torch.clamp(torch.round(x * 1000.0), -32768, 32767).to(torch.int16)
        else:
            out[c0:c0 + m] = x.to(dtype)
    return out


def corpus_cuda_range(seed, n_total, lo, hi, L=5000, dtype=None, device=None, chunk=2048):
    """Records [lo, hi) of corpus_cuda(seed, n_total, ...) without generating the rest: the generator is seeded
    per chunk of `chunk` records, so a rank of a sharded run rebuilds exactly its slice of the one corpus."""
    import torch
    dtype = dtype or torch.float32
    device = torch.device(device if device is not None else "cuda")
    out = torch.empty((max(hi - lo, 0), N_LEADS, L), dtype=dtype, device=device)
    for c0 in range((lo // chunk) * chunk, hi, chunk):
        m = min(chunk, n_total - c0)
        x = corpus_cuda(seed, m, L, dtype, device, start=c0, chunk=chunk)
        a, b = max(lo, c0), min(hi, c0 + m)
        out[a - lo:b - lo] = x[a - c0:b - c0]
    return out
