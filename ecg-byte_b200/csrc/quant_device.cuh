// Device-side sample -> symbol classification shared by the quantise and the fused
// encode kernels.  See ecgb_quantizer_create (include/ecgbyte.h) for why comparing
// against thresholds is bit-identical to tokenizer_utils.py:14-19.
#pragma once
#include "common.h"

namespace ecgb {

template <int DT> struct SampleTraits;
template <> struct SampleTraits<ECGB_F32> { using In = float;   using Thr = float;  static constexpr int kPer16B = 4; };
template <> struct SampleTraits<ECGB_F64> { using In = double;  using Thr = double; static constexpr int kPer16B = 2; };
template <> struct SampleTraits<ECGB_I16> { using In = int16_t; using Thr = float;  static constexpr int kPer16B = 8; };

// Shared-memory image of the quantiser tables.
template <typename Thr>
struct QuantSmem {
    Thr cell_thr[kCells];
    uint8_t cell_base[kCells];
};

template <typename Thr>
__device__ __forceinline__ void load_quant_smem(QuantSmem<Thr> *s, const QuantTables &t) {
    const Thr *g_thr = static_cast<const Thr *>(t.d_cell_thr);
    for (int i = threadIdx.x; i < kCells; i += blockDim.x) {
        s->cell_thr[i] = g_thr[i];
        s->cell_base[i] = t.d_cell_base[i];
    }
}

// cell index: monotone non-decreasing in the sample (every step is a monotone
// rounding operation), NaN -> cell 0.
__device__ __forceinline__ int cell_of(float sf, float lo, float scale) {
    float x = __fmul_rn(__fsub_rn(sf, lo), scale);
    x = fminf(fmaxf(x, 0.0f), (float)(kCells - 1));
    return __float2int_rz(x);
}

template <typename Thr>
__device__ __forceinline__ uint32_t classify(Thr s, float sf, float lo, float scale,
                                             const QuantSmem<Thr> *q) {
    int c = cell_of(sf, lo, scale);
    return (uint32_t)q->cell_base[c] + (s >= q->cell_thr[c] ? 1u : 0u);
}

// generic (always valid) path: count thresholds <= s.  thr has kNumThresholds entries.
template <typename Thr>
__device__ __forceinline__ uint32_t classify_search(Thr s, const Thr *thr) {
    uint32_t q = 0;
#pragma unroll
    for (int k = 0; k < kNumThresholds; k++) q += (s >= thr[k]) ? 1u : 0u;
    return q;
}

__device__ __forceinline__ float to_thr(float v, float *sf) { *sf = v; return v; }
__device__ __forceinline__ double to_thr(double v, float *sf) { *sf = __double2float_rn(v); return v; }
__device__ __forceinline__ float to_thr(int16_t v, float *sf) { *sf = (float)v; return (float)v; }

}  // namespace ecgb
