// Dataset statistics in front of the quantiser:
//   compute_global_stats  ecg_byte/utils/preprocess_utils.py:168-213
//     global_min / global_max over every stored sample      -> ecgb_minmax   (one HBM pass)
//     np.percentile(samples, 1) / (samples, 99)             -> ecgb_percentiles
// np.percentile (method 'linear') needs the order statistics around the virtual index
// (n - 1) q / 100; they are found by an 8-pass radix select over order-preserving 64-bit keys,
// the interpolation itself is done in float64 exactly as NumPy's _lerp does.
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "common.h"

namespace ecgb {

__device__ __forceinline__ unsigned long long f64_key_dev(double f) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(f);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_from_key_dev(unsigned long long k) {
    const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

template <typename T>
__global__ void __launch_bounds__(256) minmax_kernel(const T *__restrict__ in, size_t n, unsigned long long *out /*min,max,nan*/) {
    double lo = INFINITY, hi = -INFINITY;
    unsigned nan = 0;
    constexpr int V = 16 / sizeof(T);
    const size_t nv = n / V;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < nv; g += stride) {
        const uint4 raw = __ldcs(reinterpret_cast<const uint4 *>(in) + g);
        const T *e = reinterpret_cast<const T *>(&raw);
#pragma unroll
        for (int k = 0; k < V; k++) {
            const double v = (double)e[k];
            nan |= v != v;
            lo = fmin(lo, v);
            hi = fmax(hi, v);
        }
    }
    for (size_t i = nv * V + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = (double)in[i];
        nan |= v != v;
        lo = fmin(lo, v);
        hi = fmax(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        nan |= __shfl_xor_sync(0xffffffffu, nan, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&out[0], f64_key_dev(lo));
        atomicMax(&out[1], f64_key_dev(hi));
        if (nan) atomicExch(&out[2], 1ull);
    }
}

// One CTA per requested rank: the k-th smallest key (0-based) by most-significant-byte-first
// radix select; 8 passes over the samples.
__global__ void __launch_bounds__(1024) select_kernel(const double *__restrict__ x, size_t n, const unsigned long long *ranks,
                                                      double *out) {
    __shared__ unsigned long long s_hist[256];
    __shared__ unsigned long long s_prefix, s_k;
    unsigned long long prefix = 0, k = ranks[blockIdx.x];
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        const unsigned long long hi_mask = pass == 0 ? 0ull : (~0ull << (shift + 8));
        for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned long long key = f64_key_dev(x[i]);
            if ((key & hi_mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 0xFF], 1ull);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long acc = 0;
            int b = 0;
            for (; b < 256; b++) {
                if (acc + s_hist[b] > k) break;
                acc += s_hist[b];
            }
            s_prefix = prefix | ((unsigned long long)b << shift);
            s_k = k - acc;
        }
        __syncthreads();
        prefix = s_prefix;
        k = s_k;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = f64_from_key_dev(prefix);
}

}  // namespace ecgb

using namespace ecgb;

extern "C" int ecgb_minmax(const void *d_in, ecgb_dtype dtype, size_t n, double *h_min, double *h_max, int device,
                           void *stream) {
    ECGB_REQUIRE(h_min && h_max, "NULL output");
    ECGB_REQUIRE(n > 0 && d_in, "empty input");
    ECGB_REQUIRE(((uintptr_t)d_in & 15) == 0, "d_in must be 16-byte aligned");
    int rc = check_device(device);
    if (rc) return rc;
    DeviceGuard g(device);
    cudaStream_t st = as_stream(stream);
    AsyncBuf<unsigned long long> d_out;  // released on every exit path
    ECGB_CUDA(d_out.alloc(3, st));
    const unsigned long long init[3] = {~0ull, 0ull, 0ull};
    ECGB_CUDA(cudaMemcpyAsync(d_out.p, init, 24, cudaMemcpyHostToDevice, st));
    const int grid = (int)std::min<size_t>((size_t)sm_count(device) * 8, (n / 4 + 255) / 256 + 1);
    switch (dtype) {
        case ECGB_F32: minmax_kernel<float><<<grid, 256, 0, st>>>((const float *)d_in, n, d_out); break;
        case ECGB_F64: minmax_kernel<double><<<grid, 256, 0, st>>>((const double *)d_in, n, d_out); break;
        case ECGB_I16: minmax_kernel<int16_t><<<grid, 256, 0, st>>>((const int16_t *)d_in, n, d_out); break;
        default: return fail(ECGB_EINVAL, "unsupported dtype %d", (int)dtype);
    }
    ECGB_CUDA(cudaGetLastError());
    unsigned long long h[3];
    ECGB_CUDA(cudaMemcpyAsync(h, d_out.p, 24, cudaMemcpyDeviceToHost, st));
    ECGB_CUDA(cudaStreamSynchronize(st));
    auto from_key = [](unsigned long long k) {
        unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
        double f;
        std::memcpy(&f, &u, 8);
        return f;
    };
    if (h[2]) {  // np.min / np.max propagate NaN
        *h_min = *h_max = std::numeric_limits<double>::quiet_NaN();
    } else {
        *h_min = from_key(h[0]);
        *h_max = from_key(h[1]);
    }
    return ECGB_OK;
}

extern "C" int ecgb_percentiles(const double *d_samples, size_t n, const double *h_q, int nq, double *h_out, int device,
                                void *stream) {
    ECGB_REQUIRE(d_samples && h_q && h_out, "NULL argument");
    ECGB_REQUIRE(n > 0 && nq > 0 && nq <= 64, "need n > 0 and 1..64 percentiles");
    int rc = check_device(device);
    if (rc) return rc;
    DeviceGuard g(device);
    cudaStream_t st = as_stream(stream);
    // NumPy (_function_base_impl._quantile, method 'linear'):
    //   quantile = q / 100 ; virtual = (n - 1) * quantile ; previous = floor(virtual),
    //   next = previous + 1 (both the last index when virtual >= n - 1), gamma = virtual - previous
    std::vector<unsigned long long> ranks(2 * (size_t)nq);
    std::vector<double> gamma((size_t)nq);
    for (int i = 0; i < nq; i++) {
        ECGB_REQUIRE(h_q[i] >= 0.0 && h_q[i] <= 100.0, "percentiles must be in [0, 100]");
        volatile double quant = h_q[i] / 100.0;
        volatile double virt = (double)(n - 1) * quant;
        double prev = std::floor(virt);
        double nxt = prev + 1.0;
        if (virt >= (double)(n - 1)) prev = nxt = (double)(n - 1);
        if (virt < 0) prev = nxt = 0;
        ranks[2 * i] = (unsigned long long)prev;
        ranks[2 * i + 1] = (unsigned long long)nxt;
        gamma[i] = virt - std::floor(virt);
    }
    AsyncBuf<unsigned long long> d_ranks, d_mm;  // released on every exit path
    AsyncBuf<double> d_vals;
    ECGB_CUDA(d_ranks.alloc(ranks.size(), st));
    ECGB_CUDA(d_vals.alloc(ranks.size(), st));
    ECGB_CUDA(d_mm.alloc(3, st));
    const unsigned long long init[3] = {~0ull, 0ull, 0ull};
    ECGB_CUDA(cudaMemcpyAsync(d_mm.p, init, 24, cudaMemcpyHostToDevice, st));
    ECGB_CUDA(cudaMemcpyAsync(d_ranks.p, ranks.data(), ranks.size() * 8, cudaMemcpyHostToDevice, st));
    minmax_kernel<double><<<sm_count(device) * 2, 256, 0, st>>>(d_samples, n, d_mm);  // NaN detection
    select_kernel<<<2 * nq, 1024, 0, st>>>(d_samples, n, d_ranks, d_vals);
    ECGB_CUDA(cudaGetLastError());
    std::vector<double> vals(ranks.size());
    unsigned long long mm[3];
    ECGB_CUDA(cudaMemcpyAsync(vals.data(), d_vals.p, vals.size() * 8, cudaMemcpyDeviceToHost, st));
    ECGB_CUDA(cudaMemcpyAsync(mm, d_mm.p, 24, cudaMemcpyDeviceToHost, st));
    ECGB_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < nq; i++) {
        if (mm[2]) { h_out[i] = std::numeric_limits<double>::quiet_NaN(); continue; }  // NaN in, NaN out
        // NumPy _lerp: a + (b - a) * t, and b - (b - a) * (1 - t) where t >= 0.5
        volatile double a = vals[2 * i], b = vals[2 * i + 1], t = gamma[i];
        volatile double diff = b - a;
        volatile double r;
        if (t >= 0.5) { volatile double m = diff * (1.0 - t); r = b - m; } else { volatile double m = diff * t; r = a + m; }
        h_out[i] = r;
    }
    return ECGB_OK;
}
