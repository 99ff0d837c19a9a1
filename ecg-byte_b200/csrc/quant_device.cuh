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

// Shared-memory image of the quantiser cell table.
template <typename Thr>
struct QuantSmem {
    Thr cell_thr[kCells];
};

template <typename Thr>
__device__ __forceinline__ void load_quant_smem(QuantSmem<Thr> *s, const QuantTables &t) {
    const Thr *g_thr = static_cast<const Thr *>(t.d_cell_thr);
    for (int i = threadIdx.x; i < kCells; i += blockDim.x) s->cell_thr[i] = g_thr[i];
}

// cell index: monotone non-decreasing in the sample (every step is a monotone rounding
// operation); NaN and negative offsets saturate to cell 0, large ones to the last cell.
__device__ __forceinline__ uint32_t cell_of(float sf, float lo, float scale) {
    const float x = __fmul_rn(__fsub_rn(sf, lo), scale);
    return min(__float2uint_rz(x), (uint32_t)kNumThresholds);
}

template <typename Thr>
__device__ __forceinline__ uint32_t classify(Thr s, float sf, float lo, float scale,
                                             const QuantSmem<Thr> *q) {
    const uint32_t c = cell_of(sf, lo, scale);
    return c + (s >= q->cell_thr[c] ? 1u : 0u);
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
