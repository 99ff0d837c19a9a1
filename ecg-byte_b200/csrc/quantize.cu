// Q1: normalize_all (reference: ecg_byte/utils/tokenizer_utils.py:14-19) on sm_100a.
//
// The reference expression is a monotone step function of the sample, so the 25
// symbol boundaries are found once on the host by bisection over the ordered bit
// patterns of the stored type, using the float64 expression itself.  The streaming
// kernel then needs one fp32 subtract/multiply to find a pre-classification cell and
// one compare against that cell's threshold: exact, and free of the float64 divide
// that would otherwise make the kernel FP64-pipe bound instead of HBM bound.
#include <cmath>
#include <cstring>
#include <limits>
#include <new>
#include <vector>

#include "common.h"
#include "quant_device.cuh"

namespace ecgb {

// ---- the reference expression (host, float64, no contraction: built with
// -ffp-contract=off; none of these operations can fuse anyway) ----
static int ref_level(double s, double lo, double den) {
    double v = (s - lo) / den;
    if (v != v) return 0;  // NaN -> uint8 0 (x86-64 NumPy cast)
    double c = v < 0.0 ? 0.0 : v;
    c = c > 1.0 ? 1.0 : c;
    double f = std::floor(c * 26.0);
    if (f > 25.0) f = 25.0;
    return (int)f;
}

// order-preserving integer keys of float / double (total order -inf .. +inf, NaN excluded)
static uint32_t f32_key(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static float f32_from_key(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
static uint64_t f64_key(double f) {
    uint64_t u;
    std::memcpy(&u, &f, 8);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
static double f64_from_key(uint64_t k) {
    uint64_t u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double f;
    std::memcpy(&f, &u, 8);
    return f;
}

// smallest sample (in the stored type's order) whose level is >= k
static double threshold_f32(int k, double lo, double den) {
    uint32_t a = f32_key(-std::numeric_limits<float>::infinity());
    uint32_t b = f32_key(std::numeric_limits<float>::infinity());
    // level(a) = 0 < k <= 25 = level(b)
    while (b - a > 1) {
        uint32_t m = a + (b - a) / 2;
        if (ref_level((double)f32_from_key(m), lo, den) >= k) b = m; else a = m;
    }
    return (double)f32_from_key(b);
}
static double threshold_f64(int k, double lo, double den) {
    uint64_t a = f64_key(-std::numeric_limits<double>::infinity());
    uint64_t b = f64_key(std::numeric_limits<double>::infinity());
    while (b - a > 1) {
        uint64_t m = a + (b - a) / 2;
        if (ref_level(f64_from_key(m), lo, den) >= k) b = m; else a = m;
    }
    return f64_from_key(b);
}
// int16: threshold in the integer domain; +inf when no int16 value reaches level k
static double threshold_i16(int k, double lo, double den, double scale) {
    if (ref_level(32767.0 * scale, lo, den) < k) return std::numeric_limits<double>::infinity();
    if (ref_level(-32768.0 * scale, lo, den) >= k) return -32768.0;
    int a = -32768, b = 32767;  // level(a) < k <= level(b)
    while (b - a > 1) {
        int m = a + (b - a) / 2;
        if (ref_level((double)m * scale, lo, den) >= k) b = m; else a = m;
    }
    return (double)b;
}

static int host_cell(float sf, float lo, float scale) {
    volatile float d = sf - lo;
    volatile float x = d * scale;
    float y = x;
    if (!(y >= 0.0f)) y = 0.0f;  // also NaN -> 0 (fmaxf(NaN, 0) == 0)
    if (y > (float)kNumThresholds) y = (float)kNumThresholds;
    return (int)y;
}

// ---------------------------------------------------------------- kernels

template <int DT>
__global__ void __launch_bounds__(256) quantize_kernel(const typename SampleTraits<DT>::In *__restrict__ in,
                                                       size_t n, uint8_t *__restrict__ out, QuantTables tab) {
    using In = typename SampleTraits<DT>::In;
    using Thr = typename SampleTraits<DT>::Thr;
    constexpr int V = SampleTraits<DT>::kPer16B;  // samples per 16-byte load
    constexpr int NV = 16 / V;                    // loads per 16 output symbols
    __shared__ QuantSmem<Thr> qs;
    __shared__ Thr s_thr[kNumThresholds];
    load_quant_smem(&qs, tab);
    if (threadIdx.x < kNumThresholds) s_thr[threadIdx.x] = static_cast<const Thr *>(tab.d_thr)[threadIdx.x];
    __syncthreads();
    const float lo = tab.lo, scale = tab.scale;
    const bool cells = tab.exact_cells != 0;

    const size_t n16 = n / 16;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < n16; g += stride) {
        const uint4 *src = reinterpret_cast<const uint4 *>(in + g * 16);
        uint4 raw[NV];
#pragma unroll
        for (int j = 0; j < NV; j++) raw[j] = __ldcs(src + j);
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const In *e = reinterpret_cast<const In *>(&raw[j]);
#pragma unroll
            for (int k = 0; k < V; k++) {
                float sf;
                Thr s = to_thr(e[k], &sf);
                uint32_t q = cells ? classify<Thr>(s, sf, lo, scale, &qs) : classify_search<Thr>(s, s_thr);
                int idx = j * V + k;
                w[idx >> 2] |= (97u + q) << ((idx & 3) * 8);
            }
        }
        __stcs(reinterpret_cast<uint4 *>(out + g * 16), make_uint4(w[0], w[1], w[2], w[3]));
    }
    // tail (< 16 samples), one thread each
    size_t tail0 = n16 * 16;
    size_t t = tail0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        float sf;
        Thr s = to_thr(in[t], &sf);
        uint32_t q = cells ? classify<Thr>(s, sf, lo, scale, &qs) : classify_search<Thr>(s, s_thr);
        out[t] = (uint8_t)(97u + q);
    }
}

// Reference operation order on the device: float64 subtract, IEEE divide, clip,
// multiply, floor, min (tokenizer_utils.py:15-17).  __d*_rn intrinsics never contract.
template <typename In>
__global__ void __launch_bounds__(256) quantize_direct_kernel(const In *__restrict__ in, size_t n,
                                                              uint8_t *__restrict__ out, double lo, double den,
                                                              double in_scale, int scaled) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double s = (double)in[i];
        if (scaled) s = __dmul_rn(s, in_scale);
        double v = __ddiv_rn(__dsub_rn(s, lo), den);
        uint32_t q = 0;
        if (v == v) {
            double c = v < 0.0 ? 0.0 : v;
            c = c > 1.0 ? 1.0 : c;
            double f = floor(__dmul_rn(c, 26.0));
            if (f > 25.0) f = 25.0;
            q = (uint32_t)f;
        }
        out[i] = (uint8_t)(97u + q);
    }
}

static int launch_quantize(const ecgb_quantizer *q, const void *d_in, size_t n, uint8_t *d_out, cudaStream_t st) {
    if (n == 0) return ECGB_OK;
    int sms = sm_count(q->device);
    size_t want = (n / 16 + 255) / 256;
    if (want < 1) want = 1;
    size_t cap = (size_t)sms * 8;
    int grid = (int)(want < cap ? want : cap);
    // the tail needs at least ceil(15/256) = 1 block: always true
    switch (q->dtype) {
        case ECGB_F32: quantize_kernel<ECGB_F32><<<grid, 256, 0, st>>>((const float *)d_in, n, d_out, q->tab); break;
        case ECGB_F64: quantize_kernel<ECGB_F64><<<grid, 256, 0, st>>>((const double *)d_in, n, d_out, q->tab); break;
        case ECGB_I16: quantize_kernel<ECGB_I16><<<grid, 256, 0, st>>>((const int16_t *)d_in, n, d_out, q->tab); break;
        default: return fail(ECGB_EINVAL, "quantizer dtype %d cannot be quantised", (int)q->dtype);
    }
    ECGB_CUDA(cudaGetLastError());
    return ECGB_OK;
}

}  // namespace ecgb

using namespace ecgb;

extern "C" int ecgb_quantizer_create(double p1, double p99, ecgb_dtype dtype, double i16_scale, int device,
                                     ecgb_quantizer **out) {
    ECGB_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    ECGB_REQUIRE(dtype == ECGB_F32 || dtype == ECGB_F64 || dtype == ECGB_I16, "unsupported sample dtype %d", (int)dtype);
    ECGB_REQUIRE(std::isfinite(p1) && std::isfinite(p99), "percentiles must be finite");
    if (dtype == ECGB_I16) ECGB_REQUIRE(std::isfinite(i16_scale) && i16_scale > 0.0, "i16_scale must be finite and > 0");
    int rc = check_device(device);
    if (rc) return rc;
    // tokenizer_utils.py:15 -- (p1 - 0.5) and ((p99 + 0.5) - (p1 - 0.5) + 1e-6), left to right
    volatile double lo = p1 - 0.5;
    volatile double hi = p99 + 0.5;
    volatile double d0 = hi - lo;
    volatile double den = d0 + 1e-6;
    ECGB_REQUIRE(den > 0.0 && std::isfinite((double)den),
                 "quantiser denominator (p99+0.5)-(p1-0.5)+1e-6 = %g is not > 0: not a monotone quantiser", (double)den);

    ecgb_quantizer *q = new (std::nothrow) ecgb_quantizer();
    if (!q) return fail(ECGB_ENOMEM, "host allocation failed");
    q->device = device; q->dtype = dtype; q->p1 = p1; q->p99 = p99;
    q->i16_scale = dtype == ECGB_I16 ? i16_scale : 1.0;
    q->lo = lo; q->den = den; q->d_block = nullptr;
    for (int k = 1; k <= kNumThresholds; k++) {
        double t = dtype == ECGB_F32 ? threshold_f32(k, lo, den)
                 : dtype == ECGB_F64 ? threshold_f64(k, lo, den)
                                     : threshold_i16(k, lo, den, q->i16_scale);
        q->thr[k - 1] = t;
    }

    // ---- cell table: cell k must hold exactly threshold t_{k+1} ----
    const bool f64 = dtype == ECGB_F64;
    const size_t thr_sz = f64 ? 8 : 4;
    float lo_f = 0.f, scale_f = 0.f;
    int exact = 0;
    double t1 = q->thr[0], t25 = q->thr[kNumThresholds - 1];
    if (std::isfinite(t1) && std::isfinite(t25) && t25 > t1) {
        const double bin = (t25 - t1) / (double)(kNumThresholds - 1);
        scale_f = (float)(1.0 / bin);
        lo_f = (float)(t1 - 0.5 * bin);
        if (std::isfinite(scale_f) && std::isfinite(lo_f) && scale_f > 0.f) {
            exact = 1;
            for (int k = 0; k < kNumThresholds; k++)
                if (!std::isfinite(q->thr[k]) || host_cell((float)q->thr[k], lo_f, scale_f) != k) exact = 0;
        }
    }
    size_t bytes = thr_sz * kCells + thr_sz * (kNumThresholds + 2) + 64;
    std::vector<uint8_t> host(bytes, 0);
    uint8_t *h_cell_thr = host.data();
    uint8_t *h_thr = host.data() + thr_sz * kCells;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    auto put = [&](uint8_t *base, int i, double v) {
        if (f64) { std::memcpy(base + 8 * i, &v, 8); } else { float f = (float)v; std::memcpy(base + 4 * i, &f, 4); }
    };
    for (int c = 0; c < kCells; c++) put(h_cell_thr, c, c < kNumThresholds ? q->thr[c] : nan);
    for (int k = 0; k < kNumThresholds; k++) put(h_thr, k, q->thr[k]);
    put(h_thr, kNumThresholds, nan);
    put(h_thr, kNumThresholds + 1, nan);

    DeviceGuard g(device);
    cudaError_t e = cudaMalloc(&q->d_block, bytes);
    if (e != cudaSuccess) { delete q; return fail(ECGB_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
    e = cudaMemcpy(q->d_block, host.data(), bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(q->d_block); delete q; return fail(ECGB_ECUDA, "cudaMemcpy failed: %s", cudaGetErrorString(e)); }
    uint8_t *d = static_cast<uint8_t *>(q->d_block);
    q->tab.lo = lo_f; q->tab.scale = scale_f; q->tab.exact_cells = exact;
    q->tab.d_cell_thr = d;
    q->tab.d_thr = d + thr_sz * kCells;
    *out = q;
    return ECGB_OK;
}

extern "C" int ecgb_quantizer_destroy(ecgb_quantizer *q) {
    if (!q) return ECGB_OK;
    if (q->d_block) { DeviceGuard g(q->device); cudaFree(q->d_block); }
    delete q;
    return ECGB_OK;
}

extern "C" int ecgb_quantizer_thresholds(const ecgb_quantizer *q, double h_thr_out[25]) {
    ECGB_REQUIRE(q && h_thr_out, "NULL argument");
    for (int k = 0; k < kNumThresholds; k++) h_thr_out[k] = q->thr[k];
    return ECGB_OK;
}

extern "C" int ecgb_quantize(const ecgb_quantizer *q, const void *d_in, size_t n, uint8_t *d_out, void *stream) {
    ECGB_REQUIRE(q, "quantizer is NULL");
    ECGB_REQUIRE(n == 0 || (d_in && d_out), "NULL buffer");
    ECGB_REQUIRE(((uintptr_t)d_in & 15) == 0 && ((uintptr_t)d_out & 15) == 0, "d_in / d_out must be 16-byte aligned");
    DeviceGuard g(q->device);
    return launch_quantize(q, d_in, n, d_out, as_stream(stream));
}

extern "C" int ecgb_quantize_direct(const ecgb_quantizer *q, const void *d_in, size_t n, uint8_t *d_out, void *stream) {
    ECGB_REQUIRE(q, "quantizer is NULL");
    ECGB_REQUIRE(n == 0 || (d_in && d_out), "NULL buffer");
    if (n == 0) return ECGB_OK;
    DeviceGuard g(q->device);
    cudaStream_t st = as_stream(stream);
    int grid = sm_count(q->device) * 8;
    switch (q->dtype) {
        case ECGB_F32: quantize_direct_kernel<float><<<grid, 256, 0, st>>>((const float *)d_in, n, d_out, q->lo, q->den, 1.0, 0); break;
        case ECGB_F64: quantize_direct_kernel<double><<<grid, 256, 0, st>>>((const double *)d_in, n, d_out, q->lo, q->den, 1.0, 0); break;
        case ECGB_I16: quantize_direct_kernel<int16_t><<<grid, 256, 0, st>>>((const int16_t *)d_in, n, d_out, q->lo, q->den, q->i16_scale, 1); break;
        default: return fail(ECGB_EINVAL, "bad dtype");
    }
    ECGB_CUDA(cudaGetLastError());
    return ECGB_OK;
}

extern "C" int ecgb_quantize_host(const ecgb_quantizer *q, const void *h_in, size_t n, uint8_t *h_out) {
    ECGB_REQUIRE(q, "quantizer is NULL");
    ECGB_REQUIRE(n == 0 || (h_in && h_out), "NULL buffer");
    if (n == 0) return ECGB_OK;
    DeviceGuard g(q->device);
    size_t es = q->dtype == ECGB_F64 ? 8 : q->dtype == ECGB_F32 ? 4 : 2;
    void *d_in = nullptr; uint8_t *d_out = nullptr;
    ECGB_CUDA(cudaMalloc(&d_in, n * es));
    cudaError_t e = cudaMalloc(&d_out, n);
    if (e != cudaSuccess) { cudaFree(d_in); return fail(ECGB_ENOMEM, "cudaMalloc failed: %s", cudaGetErrorString(e)); }
    int rc = ECGB_OK;
    e = cudaMemcpy(d_in, h_in, n * es, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) rc = launch_quantize(q, d_in, n, d_out, 0);
    if (e == cudaSuccess && rc == ECGB_OK) e = cudaMemcpy(h_out, d_out, n, cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) return fail(ECGB_ECUDA, "copy failed: %s", cudaGetErrorString(e));
    return rc;
}
