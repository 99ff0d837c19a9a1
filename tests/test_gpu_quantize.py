"""K1 parity: CUDA quantiser vs the CPU oracle (and vs the float64 expression run on
the device), bit-exact, on every stored dtype and on the edge values."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

DT = {np.float32: "float32", np.float64: "float64", np.int16: "int16"}


def _edge_values(dtype, thr):
    if dtype == np.int16:
        base = np.array([-32768, -1, 0, 1, 32767], np.int16)
        near = np.clip(np.concatenate([thr[np.isfinite(thr)] + d for d in (-1, 0, 1)]), -32768, 32767).astype(np.int16)
        return np.concatenate([base, near])
    fin = thr[np.isfinite(thr)].astype(dtype)
    near = np.concatenate([fin, np.nextafter(fin, dtype(-np.inf)), np.nextafter(fin, dtype(np.inf))])
    special = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1e30, -1e30, 1e-40, np.finfo(dtype).max, np.finfo(dtype).min],
                       dtype)
    return np.concatenate([special, near])


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int16])
@pytest.mark.parametrize("pct", [(-0.16, 0.9), (-1.5, 2.25), (0.0, 0.0), (-3000.0, 4000.0)])
def test_quantize_matches_oracle(oracle, dtype, pct):
    from ecgbyte.api import Quantizer
    p = {"percentile_1": pct[0], "percentile_99": pct[1]}
    scale = 1e-3 if abs(pct[0]) < 100 else 1.0
    q = Quantizer(p, dtype=getattr(torch, DT[dtype]), i16_scale=scale)
    thr = q.thresholds()
    rng = np.random.default_rng(1)
    lo, hi = pct[0] - 1.0, pct[1] + 1.0
    if dtype == np.int16:
        x = rng.integers(-32768, 32768, size=100003).astype(np.int16)
    else:
        x = rng.uniform(lo - 0.3 * (hi - lo), hi + 0.3 * (hi - lo), size=100003).astype(dtype)
    x = np.concatenate([x, _edge_values(dtype, thr)])
    want = oracle.quantize(x, pct[0], pct[1], scale)
    xd = torch.from_numpy(x).cuda()
    got = q.quantize(xd).cpu().numpy()
    np.testing.assert_array_equal(got, want)
    direct = q.quantize(xd, direct=True).cpu().numpy()
    np.testing.assert_array_equal(direct, want)
    np.testing.assert_array_equal(q.quantize_host(x), want)


def test_quantize_rejects_non_monotone():
    from ecgbyte.api import Quantizer
    with pytest.raises(ValueError):
        Quantizer({"percentile_1": 5.0, "percentile_99": 0.0})
    with pytest.raises(ValueError):
        Quantizer({"percentile_1": float("nan"), "percentile_99": 0.0})


def test_quantize_empty_and_ragged_sizes(oracle):
    from ecgbyte.api import Quantizer
    p = {"percentile_1": -0.2, "percentile_99": 0.8}
    q = Quantizer(p)
    rng = np.random.default_rng(3)
    for n in (0, 1, 15, 16, 17, 31, 4097):
        x = rng.normal(0.2, 0.5, size=n).astype(np.float32)
        got = q.quantize(torch.from_numpy(x).cuda()).cpu().numpy()
        np.testing.assert_array_equal(got, oracle.quantize(x, -0.2, 0.8))


def test_normalize_all_mirror(oracle, small_corpus):
    """tokenizer_utils.normalize_all keeps the reference's return types."""
    from ecgbyte import tokenizer_utils as tu
    x, pct = small_corpus
    clipped, sym = tu.normalize_all(x[0], pct)
    assert sym.dtype == np.dtype("<U1") and sym.shape == x[0].shape
    want = oracle.quantize(x[0], pct["percentile_1"], pct["percentile_99"])
    assert "".join(sym.flatten()) == want.tobytes().decode()
    ref = np.clip((x[0] - (pct["percentile_1"] - 0.5)) /
                  ((pct["percentile_99"] + 0.5) - (pct["percentile_1"] - 0.5) + 1e-6), 0, 1)
    np.testing.assert_array_equal(clipped, ref)
    back = tu.reverse_normalize_all(sym, pct)
    assert back.shape == x[0].shape
