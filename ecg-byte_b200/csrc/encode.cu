// E2: encode_text (reference: rust_bpe/src/lib.rs:149-193) fused with Q1
// (tokenizer_utils.py:14-19) for sm_100a.
//
// Semantics: greedy longest match over the trie -- at each token start walk as far
// as symbols match, remember the longest terminal seen, emit it and restart right
// after it (lib.rs:163-190).  This is NOT rank-ordered BPE merging.
//
// Mapping (one walker = one thread = one record, 1 CTA of up to 1024 walkers per SM):
//   * records are independent and 100k-1M of them are in flight, so the
//     record axis alone fills the chip: no speculation, no redundant trie steps, no
//     inter-thread synchronisation after the tables are staged;
//   * the trie (8-byte bitmap nodes, ~90 KB for 5,000 merges) is staged once per CTA in
//     shared memory; one trie step = one LDS.64 + popc;
//   * each walker streams its own record from HBM with 128-bit loads, one 32-byte sector
//     (8 fp32 samples) per request, prefetched one group ahead in registers; samples are
//     quantised by threshold classification as they arrive and kept as a 16-symbol
//     register window, so symbols never touch HBM or shared memory;
//   * a failed walk re-reads the symbols after the emitted token from the window
//     (99.3 % of restarts reach back <= 8 symbols on ECG data); a longer reach-back
//     re-primes the window from L1/L2.
// HBM traffic is therefore the algorithmic minimum: samples once, tokens once.
#include <algorithm>
#include <cstring>

#include "common.h"
#include "quant_device.cuh"

namespace ecgb {

struct EncArgs {
    const void *in;            // samples (or text bytes), all records back to back
    size_t n_total;            // total samples in `in`
    size_t n_rec, rec_len;
    const uint64_t *offsets;   // optional [n_rec + 1]
    int32_t *tokens;
    size_t out_stride;
    int32_t *lens;
    const uint2 *nodes;
    uint32_t n_nodes, smem_nodes;
    const uint8_t *cls;
    QuantTables qt;
};

constexpr uint32_t kNoTok = 0xFFFFu;
constexpr uint32_t kNoClass = 31u;

// ---- 8-sample groups: the unit a walker loads, quantises and slides by ----
template <int DT> struct RawGroup;
template <> struct RawGroup<ECGB_F32> { uint4 v[2]; };
template <> struct RawGroup<ECGB_F64> { uint4 v[4]; };
template <> struct RawGroup<ECGB_I16> { uint4 v[1]; };
template <> struct RawGroup<ECGB_U8>  { uint2 v[1]; };

template <int DT> struct ElemOf { using T = typename SampleTraits<DT>::In; };
template <> struct ElemOf<ECGB_U8> { using T = uint8_t; };

template <int DT>
__device__ __forceinline__ void load_group(RawGroup<DT> &r, const void *base, size_t g, size_t n_total) {
    using T = typename ElemOf<DT>::T;
    const T *p = static_cast<const T *>(base) + g;
    if (g + 8 <= n_total) {
        if constexpr (DT == ECGB_U8) {
            r.v[0] = __ldg(reinterpret_cast<const uint2 *>(p));
        } else {
            constexpr int NV = sizeof(T) * 8 / 16;
#pragma unroll
            for (int j = 0; j < NV; j++) r.v[j] = __ldg(reinterpret_cast<const uint4 *>(p) + j);
        }
    } else {  // ragged end of the buffer: element-wise, zero filled
        T tmp[8];
#pragma unroll
        for (int k = 0; k < 8; k++) tmp[k] = (g + k < n_total) ? p[k] : T(0);
        memcpy(&r, tmp, sizeof(r));
    }
}

// 8 samples -> 8 symbol classes packed one per byte (x = samples 0..3, y = 4..7).
template <int DT, bool CELLS>
__device__ __forceinline__ uint2 quantize_group(const RawGroup<DT> &r, const void *qsmem, const void *thr_smem,
                                                float lo, float scale) {
    if constexpr (DT == ECGB_U8) {
        return r.v[0];  // text bytes; class lookup happens per step
    } else {
        using T = typename SampleTraits<DT>::In;
        using Thr = typename SampleTraits<DT>::Thr;
        const T *e = reinterpret_cast<const T *>(&r);
        uint32_t w[2] = {0, 0};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float sf;
            Thr s = to_thr(e[k], &sf);
            uint32_t q = CELLS ? classify<Thr>(s, sf, lo, scale, static_cast<const QuantSmem<Thr> *>(qsmem))
                               : classify_search<Thr>(s, static_cast<const Thr *>(thr_smem));
            w[k >> 2] |= q << ((k & 3) * 8);
        }
        return make_uint2(w[0], w[1]);
    }
}

template <int DT> struct ThrOf { using T = typename SampleTraits<DT>::Thr; };
template <> struct ThrOf<ECGB_U8> { using T = float; };

template <int DT, bool CELLS>
__global__ void __launch_bounds__(1024, 1) encode_kernel(EncArgs a) {
    using Thr = typename ThrOf<DT>::T;
    extern __shared__ __align__(16) uint8_t smem[];
    uint2 *s_nodes = reinterpret_cast<uint2 *>(smem);
    const size_t nodes_bytes = ((size_t)a.smem_nodes * 8 + 15) & ~(size_t)15;
    uint8_t *s_aux = smem + nodes_bytes;
    // aux region: quantiser tables (sample dtypes) or the byte->class table (text)
    QuantSmem<Thr> *qs = reinterpret_cast<QuantSmem<Thr> *>(s_aux);
    Thr *s_thr = reinterpret_cast<Thr *>(s_aux + sizeof(QuantSmem<Thr>));
    uint8_t *s_cls = s_aux;

    for (uint32_t i = threadIdx.x; i < a.smem_nodes; i += blockDim.x) s_nodes[i] = a.nodes[i];
    if constexpr (DT == ECGB_U8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_cls[i] = a.cls[i];
    } else {
        load_quant_smem(qs, a.qt);
        if (threadIdx.x < kNumThresholds) s_thr[threadIdx.x] = static_cast<const Thr *>(a.qt.d_thr)[threadIdx.x];
    }
    __syncthreads();

    const float qlo = a.qt.lo, qscale = a.qt.scale;
    const uint32_t S = a.smem_nodes;
    // explicit offsets: the buffer ends where the last record ends
    const size_t n_total = a.offsets ? (size_t)a.offsets[a.n_rec] : a.n_total;
    const uint2 root = s_nodes[0];
    const uint32_t root_mask = root.x, root_base = root.y >> 16;

    // contiguous, even split of the records over the CTAs
    const size_t r_lo = (size_t)(((unsigned __int128)a.n_rec * blockIdx.x) / gridDim.x);
    const size_t r_hi = (size_t)(((unsigned __int128)a.n_rec * (blockIdx.x + 1)) / gridDim.x);

    for (size_t r = r_lo + threadIdx.x; r < r_hi; r += blockDim.x) {
        const size_t rs = a.offsets ? (size_t)a.offsets[r] : r * a.rec_len;
        const size_t re = a.offsets ? (size_t)a.offsets[r + 1] : rs + a.rec_len;
        int32_t *outp = a.tokens + r * a.out_stride;
        uint32_t cnt = 0;

        size_t wb;       // window covers symbols [wb, wb + 16), wb % 8 == 0 (global sample index)
        uint2 w0, w1;    // 8 symbols each
        RawGroup<DT> nxt;  // group [wb + 16, wb + 24) in flight
        auto prime = [&](size_t at) {
            wb = at & ~(size_t)7;
            RawGroup<DT> g0, g1;
            load_group<DT>(g0, a.in, wb, n_total);
            load_group<DT>(g1, a.in, wb + 8, n_total);
            load_group<DT>(nxt, a.in, wb + 16, n_total);
            w0 = quantize_group<DT, CELLS>(g0, qs, s_thr, qlo, qscale);
            w1 = quantize_group<DT, CELLS>(g1, qs, s_thr, qlo, qscale);
        };
        prime(rs);

        size_t pos = rs, start = rs;
        uint32_t mask = root_mask, base = root_base, depth = 0, mlen = 0, mid = 0;
        for (;;) {
            if (pos >= wb + 16) {  // slide by one group; the next one is already in registers
                w0 = w1;
                w1 = quantize_group<DT, CELLS>(nxt, qs, s_thr, qlo, qscale);
                wb += 8;
                load_group<DT>(nxt, a.in, wb + 16, n_total);
            }
            const uint32_t off = (uint32_t)(pos - wb);
            const uint32_t lo32 = (off & 8) ? w1.x : w0.x, hi32 = (off & 8) ? w1.y : w0.y;
            const uint32_t byte = __byte_perm(lo32, hi32, off & 7) & 0xffu;
            uint32_t c = byte;
            if constexpr (DT == ECGB_U8) c = s_cls[byte];
            const bool have = pos < re;
            const bool ok = have && c < kNoClass && ((mask >> c) & 1u);
            if (ok) {
                const uint32_t idx = base + __popc(mask & ((1u << c) - 1u));
                const uint2 nd = idx < S ? s_nodes[idx] : __ldg(a.nodes + idx);
                mask = nd.x;
                base = nd.y >> 16;
                const uint32_t tok = nd.y & 0xFFFFu;
                pos++;
                depth++;
                if (tok != kNoTok) { mlen = depth; mid = tok; }
            } else if (depth == 0) {
                if (!have) break;  // record exhausted at a token boundary
                // a byte that occurs in no merge: its own single-byte token (lib.rs:155-157)
                if (cnt < a.out_stride) outp[cnt] = (int32_t)byte;
                cnt++;
                pos++;
                start = pos;
            } else {  // walk ended: emit the longest terminal, resume right after it
                if (cnt < a.out_stride) outp[cnt] = (int32_t)mid;
                cnt++;
                start += mlen;
                if (start < wb) prime(start);
                pos = start;
                depth = 0;
                mlen = 0;
                mask = root_mask;
                base = root_base;
            }
        }
        a.lens[r] = (int32_t)cnt;
    }
}

// ---- wide (any byte alphabet) vocabularies: 40-byte nodes in global memory / L1 ----
__global__ void __launch_bounds__(256) encode_wide_kernel(const uint8_t *__restrict__ text, size_t n_rec, size_t rec_len,
                                                          const uint64_t *__restrict__ offsets, int32_t *tokens,
                                                          size_t out_stride, int32_t *lens,
                                                          const uint32_t *__restrict__ nodes) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += stride) {
        const size_t rs = offsets ? (size_t)offsets[r] : r * rec_len;
        const size_t re = offsets ? (size_t)offsets[r + 1] : rs + rec_len;
        int32_t *outp = tokens + r * out_stride;
        uint32_t cnt = 0;
        size_t start = rs;
        while (start < re) {
            uint32_t node = 0, depth = 0, mlen = 0, mid = 0;
            for (size_t pos = start; pos < re; pos++) {
                const uint32_t b = text[pos];
                const uint32_t *nd = nodes + (size_t)node * 10;
                const uint32_t word = __ldg(nd + (b >> 5));
                if (!((word >> (b & 31)) & 1u)) break;
                uint32_t rank = __popc(word & ((1u << (b & 31)) - 1u));
                for (uint32_t k = 0; k < (b >> 5); k++) rank += __popc(__ldg(nd + k));
                node = __ldg(nd + 8) + rank;
                depth++;
                const uint32_t tok = __ldg(nodes + (size_t)node * 10 + 9);
                if (tok != 0xFFFFFFFFu) { mlen = depth; mid = tok; }
            }
            // every byte is a child of the root, so mlen >= 1
            if (cnt < out_stride) outp[cnt] = (int32_t)mid;
            cnt++;
            start += mlen;
        }
        lens[r] = (int32_t)cnt;
    }
}

template <int DT>
static int launch_encode_t(const EncArgs &a, int exact_cells, int device, cudaStream_t st) {
    using Thr = typename ThrOf<DT>::T;
    int sms = sm_count(device);
    int smem_max = 0;
    ECGB_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    const size_t aux = DT == ECGB_U8 ? 256 : sizeof(QuantSmem<Thr>) + sizeof(Thr) * 32;
    EncArgs args = a;
    size_t budget = (size_t)smem_max > aux + 1024 ? (size_t)smem_max - aux - 1024 : 0;
    args.smem_nodes = (uint32_t)std::min<size_t>(a.n_nodes, budget / 8);
    if (args.smem_nodes < 1) return fail(ECGB_EUNSUPPORTED, "device shared memory too small for the trie root");
    size_t smem = (((size_t)args.smem_nodes * 8 + 15) & ~(size_t)15) + aux;

    // one walker per record; CTAs get equal contiguous record ranges
    size_t grid = std::min<size_t>((size_t)sms, (a.n_rec + 31) / 32);
    if (grid < 1) grid = 1;
    size_t per_cta = (a.n_rec + grid - 1) / grid;
    int block = (int)std::min<size_t>(1024, ((per_cta + 31) / 32) * 32);
    auto kern = exact_cells ? encode_kernel<DT, true> : encode_kernel<DT, false>;
    ECGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, block, smem, st>>>(args);
    ECGB_CUDA(cudaGetLastError());
    return ECGB_OK;
}

static int launch_encode(int dt, const EncArgs &a, int exact_cells, int device, cudaStream_t st) {
    switch (dt) {
        case ECGB_F32: return launch_encode_t<ECGB_F32>(a, exact_cells, device, st);
        case ECGB_F64: return launch_encode_t<ECGB_F64>(a, exact_cells, device, st);
        case ECGB_I16: return launch_encode_t<ECGB_I16>(a, exact_cells, device, st);
        case ECGB_U8: return launch_encode_t<ECGB_U8>(a, 1, device, st);
    }
    return fail(ECGB_EINVAL, "bad dtype %d", dt);
}

}  // namespace ecgb

using namespace ecgb;

extern "C" int ecgb_encode_symbols(const ecgb_vocab *v, const uint8_t *d_sym, size_t n_rec, size_t rec_len,
                                   const uint64_t *d_offsets, int32_t *d_tokens, size_t out_stride, int32_t *d_len,
                                   void *stream) {
    ECGB_REQUIRE(v, "vocab is NULL");
    if (n_rec == 0) return ECGB_OK;
    ECGB_REQUIRE(d_len && (d_tokens || out_stride == 0), "NULL output buffer");
    ECGB_REQUIRE(d_sym || (rec_len == 0 && !d_offsets), "d_sym is NULL");
    ECGB_REQUIRE(((uintptr_t)d_sym & 7) == 0, "d_sym must be 8-byte aligned");
    const VocabView *vv = ecgb_vocab_view(v);
    int device = ecgb_vocab_device(v);
    DeviceGuard g(device);
    cudaStream_t st = as_stream(stream);
    if (!vv->compact) {
        int grid = (int)std::min<size_t>((size_t)sm_count(device) * 8, (n_rec + 255) / 256);
        encode_wide_kernel<<<grid, 256, 0, st>>>(d_sym, n_rec, rec_len, d_offsets, d_tokens, out_stride, d_len, vv->d_wide);
        ECGB_CUDA(cudaGetLastError());
        return ECGB_OK;
    }
    EncArgs a{};
    a.in = d_sym;
    a.n_total = n_rec * rec_len;  // with explicit offsets the kernel uses offsets[n_rec]
    a.n_rec = n_rec; a.rec_len = rec_len; a.offsets = d_offsets;
    a.tokens = d_tokens; a.out_stride = out_stride; a.lens = d_len;
    a.nodes = vv->d_nodes; a.n_nodes = vv->n_nodes; a.cls = vv->d_cls;
    return launch_encode(ECGB_U8, a, 1, device, st);
}

extern "C" int ecgb_encode_batch(const ecgb_vocab *v, const ecgb_quantizer *q, const void *d_in, size_t n_rec,
                                 size_t rec_len, int32_t *d_tokens, size_t out_stride, int32_t *d_len, void *stream) {
    ECGB_REQUIRE(v && q, "vocab / quantizer is NULL");
    if (n_rec == 0) return ECGB_OK;
    ECGB_REQUIRE(d_in && d_len && (d_tokens || out_stride == 0), "NULL buffer");
    ECGB_REQUIRE(((uintptr_t)d_in & 15) == 0, "d_in must be 16-byte aligned");
    const VocabView *vv = ecgb_vocab_view(v);
    int device = ecgb_vocab_device(v);
    ECGB_REQUIRE(device == q->device, "vocab (device %d) and quantizer (device %d) live on different devices", device, q->device);
    DeviceGuard g(device);
    cudaStream_t st = as_stream(stream);
    if (!vv->compact) {
        // wide vocabularies: quantise to a temporary symbol buffer, then the wide walker
        uint8_t *d_sym = nullptr;
        ECGB_CUDA(cudaMallocAsync((void **)&d_sym, n_rec * rec_len, st));
        int rc = ecgb_quantize(q, d_in, n_rec * rec_len, d_sym, stream);
        if (rc == ECGB_OK) rc = ecgb_encode_symbols(v, d_sym, n_rec, rec_len, nullptr, d_tokens, out_stride, d_len, stream);
        cudaFreeAsync(d_sym, st);
        return rc;
    }
    EncArgs a{};
    a.in = d_in; a.n_total = n_rec * rec_len;
    a.n_rec = n_rec; a.rec_len = rec_len; a.offsets = nullptr;
    a.tokens = d_tokens; a.out_stride = out_stride; a.lens = d_len;
    a.nodes = vv->d_nodes; a.n_nodes = vv->n_nodes; a.cls = vv->d_cls;
    a.qt = q->tab;
    return launch_encode((int)q->dtype, a, q->tab.exact_cells, device, st);
}

extern "C" int ecgb_encode_text_host(const ecgb_vocab *v, const uint8_t *h_text, size_t n, uint32_t *h_out, size_t cap,
                                     size_t *n_out) {
    ECGB_REQUIRE(v && n_out, "NULL argument");
    *n_out = 0;
    if (n == 0) return ECGB_OK;
    ECGB_REQUIRE(h_text, "h_text is NULL");
    int device = ecgb_vocab_device(v);
    DeviceGuard g(device);
    uint8_t *d_text = nullptr; int32_t *d_tok = nullptr; int32_t *d_len = nullptr;
    const size_t stride = std::min(cap, n);
    cudaError_t e = cudaMalloc((void **)&d_text, n + 64);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_tok, std::max<size_t>(stride, 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_len, 4);
    int rc = ECGB_OK;
    int32_t len = 0;
    if (e == cudaSuccess) e = cudaMemcpy(d_text, h_text, n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) rc = ecgb_encode_symbols(v, d_text, 1, n, nullptr, d_tok, stride, d_len, nullptr);
    if (e == cudaSuccess && rc == ECGB_OK) e = cudaMemcpy(&len, d_len, 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == ECGB_OK && h_out && stride)
        e = cudaMemcpy(h_out, d_tok, std::min<size_t>((size_t)len, stride) * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_text); cudaFree(d_tok); cudaFree(d_len);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? ECGB_ENOMEM : ECGB_ECUDA, "encode_text_host: %s", cudaGetErrorString(e));
    if (rc) return rc;
    *n_out = (size_t)len;
    if ((size_t)len > cap) return fail(ECGB_ECAPACITY, "output capacity %zu < %d tokens", cap, len);
    return ECGB_OK;
}

extern "C" int ecgb_encode_batch_host(const ecgb_vocab *v, const ecgb_quantizer *q, const void *h_in, size_t n_rec,
                                      size_t rec_len, int32_t *h_tokens, size_t out_stride, int32_t *h_len) {
    ECGB_REQUIRE(v && q, "vocab / quantizer is NULL");
    if (n_rec == 0) return ECGB_OK;
    ECGB_REQUIRE(h_in && h_len && (h_tokens || out_stride == 0), "NULL buffer");
    int device = ecgb_vocab_device(v);
    DeviceGuard g(device);
    const size_t es = q->dtype == ECGB_F64 ? 8 : q->dtype == ECGB_F32 ? 4 : 2;
    const size_t in_bytes = n_rec * rec_len * es, tok_bytes = n_rec * out_stride * 4;
    void *d_in = nullptr; int32_t *d_tok = nullptr; int32_t *d_len = nullptr;
    cudaError_t e = cudaMalloc(&d_in, in_bytes + 64);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_tok, std::max<size_t>(tok_bytes, 4));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_len, n_rec * 4);
    int rc = ECGB_OK;
    if (e == cudaSuccess) e = cudaMemcpy(d_in, h_in, in_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) rc = ecgb_encode_batch(v, q, d_in, n_rec, rec_len, d_tok, out_stride, d_len, nullptr);
    if (e == cudaSuccess && rc == ECGB_OK && tok_bytes) e = cudaMemcpy(h_tokens, d_tok, tok_bytes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == ECGB_OK) e = cudaMemcpy(h_len, d_len, n_rec * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_tok); cudaFree(d_len);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? ECGB_ENOMEM : ECGB_ECUDA, "encode_batch_host: %s", cudaGetErrorString(e));
    return rc;
}
