// The rows either side of the encoder:
//   decode_text            ecg_byte/utils/tokenizer_utils.py:75-77     -> ecgb_decode_symbols
//   reverse_normalize_all  ecg_byte/utils/tokenizer_utils.py:22-28     -> ecgb_dequantize
//   ECGTokenDataset: signal_k -> LLM id (data_loader.py:80), truncate / left-pad / labels /
//   attention mask / position ids (data_loader.py:26-31, 101-132)      -> ecgb_pack_training
//   expand_attention       ecg_byte/runners/interpret.py:106-111       -> ecgb_expand_attention
//   analyze_token_distribution (Counter over encoded ids) tokenizer_utils.py:30-54 -> ecgb_token_histogram
#include <algorithm>

#include "common.h"

namespace ecgb {

// One CTA per record: prefix sum of the token lengths, then every token copies its bytes
// (decode_text) or repeats its attention value once per symbol (ATTN: expand_attention).
template <bool ATTN, class Out>
__global__ void __launch_bounds__(256) decode_kernel(const int32_t *__restrict__ tokens, size_t in_stride,
                                                     const int32_t *__restrict__ lens, Out *__restrict__ sym,
                                                     size_t sym_stride, int32_t *__restrict__ sym_len,
                                                     const uint8_t *__restrict__ dec_sym,
                                                     const uint32_t *__restrict__ dec_off, uint32_t dec_ids, int *bad,
                                                     const float *__restrict__ attn) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_carry;
    const size_t r = blockIdx.x;
    const int32_t *tok = tokens + r * in_stride;
    Out *out = sym + r * sym_stride;
    const uint32_t n = (uint32_t)min((long long)max(lens[r], 0), (long long)in_stride);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        uint32_t o = 0, l = 0;
        if (i < n) {
            const uint32_t t = (uint32_t)tok[i];
            if (t < dec_ids) { o = dec_off[t]; l = dec_off[t + 1] - o; }
            if (l == 0) atomicExch(bad, 1);  // unknown token id
        }
        uint32_t incl = l;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t woff = 0, total = 0;
        for (int w = 0; w < 8; w++) { if (w < warp) woff += s_warp[w]; total += s_warp[w]; }
        const uint32_t dst = s_carry + woff + incl - l;
        if constexpr (ATTN) {
            const float av = i < n ? attn[r * in_stride + i] : 0.f;
            for (uint32_t k = 0; k < l; k++)
                if (dst + k < sym_stride) out[dst + k] = av;
        } else {
            for (uint32_t k = 0; k < l; k++)
                if (dst + k < sym_stride) out[dst + k] = dec_sym[o + k];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) sym_len[r] = (int32_t)s_carry;
}

// tokenizer_utils.py:25-27: (symbol index / 25) * (max - min) + min, float64, no contraction
__global__ void __launch_bounds__(256) dequantize_kernel(const uint8_t *__restrict__ sym, size_t n, double *__restrict__ out,
                                                         double min_v, double span) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double q = (double)((int)sym[i] - 97);
        out[i] = __dadd_rn(__dmul_rn(__ddiv_rn(q, 25.0), span), min_v);
    }
}

struct PackCfg {
    long long pad_id, bos_id, eos_id, sig_start_id, sig_end_id;
    uint32_t pad_to_max;
};

// data_loader.py:80 + 101-132 for one sample per CTA.  Output row length P = pad_to_max + 4:
//   [pad]*k + [bos, sig_start] + signal[:available] + [sig_end] + question + answer + [eos]
__global__ void __launch_bounds__(256) pack_kernel(const int32_t *__restrict__ tokens, size_t in_stride,
                                                   const int32_t *__restrict__ lens, const long long *__restrict__ lut,
                                                   uint32_t lut_size, const long long *__restrict__ text,
                                                   const unsigned long long *__restrict__ text_off,
                                                   const int32_t *__restrict__ q_len, PackCfg cfg,
                                                   long long *__restrict__ input_ids, float *__restrict__ attn,
                                                   long long *__restrict__ labels, long long *__restrict__ pos_ids,
                                                   int32_t *__restrict__ status) {
    __shared__ int s_warp[8];
    __shared__ int s_carry;
    const size_t r = blockIdx.x;
    const int P = (int)cfg.pad_to_max + 4;
    const int32_t *tok = tokens + r * in_stride;
    const long long *txt = text + text_off[r];
    const int qa = (int)(text_off[r + 1] - text_off[r]);
    const int nq = q_len[r];
    const int nsig_all = min(max(lens[r], 0), (int)in_stride);
    const int avail = (int)cfg.pad_to_max - qa;  // data_loader.py:103-104
    long long *ids = input_ids + r * P;
    float *am = attn + r * P;
    long long *lab = labels + r * P;
    long long *pid = pos_ids + r * P;
    if (avail < 0 || nq < 0 || nq > qa) {
        // question + answer longer than pad_to_max: the reference's length assert fails (data_loader.py:123)
        for (int i = threadIdx.x; i < P; i += blockDim.x) { ids[i] = cfg.pad_id; am[i] = 0.f; lab[i] = -100; pid[i] = 0; }
        if (threadIdx.x == 0) status[r] = 1;
        return;
    }
    const int nsig = min(nsig_all, avail);  // truncation, data_loader.py:106-107
    const int npad = avail - nsig;          // left padding, data_loader.py:108-109
    const int sig_block = npad + 2 + nsig + 1;  // [pad]*npad + bos + sig_start + signal + sig_end
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < P; base += blockDim.x) {
        const int i = base + threadIdx.x;
        long long v = cfg.pad_id, lb = -100;
        if (i < P) {
            if (i < npad) v = cfg.pad_id;
            else if (i == npad) v = cfg.bos_id;
            else if (i == npad + 1) v = cfg.sig_start_id;
            else if (i < npad + 2 + nsig) {
                const uint32_t t = (uint32_t)tok[i - npad - 2];
                v = t < lut_size ? lut[t] : cfg.pad_id;  // 'signal_{id}' -> LLM id
            } else if (i == npad + 2 + nsig) v = cfg.sig_end_id;
            else if (i < sig_block + qa) v = txt[i - sig_block];
            else v = cfg.eos_id;
            // labels: -100 over the signal block and the question, then answer + eos (data_loader.py:115)
            if (i >= sig_block + nq) lb = v;
        }
        // position ids: cumsum(mask) - 1, zero where padded (data_loader.py:26-31)
        const int m = (i < P && v != cfg.pad_id) ? 1 : 0;
        int incl = m;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int woff = 0, total = 0;
        for (int w = 0; w < 8; w++) { if (w < warp) woff += s_warp[w]; total += s_warp[w]; }
        if (i < P) {
            ids[i] = v;
            am[i] = m ? 1.f : 0.f;
            lab[i] = lb;
            pid[i] = m ? (long long)(s_carry + woff + incl - 1) : 0;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) status[r] = 0;
}

}  // namespace ecgb

namespace ecgb {
// Counter over the encoded ids of a batch (tokenizer_utils.py:44-49): block-private counts in shared
// memory while the id space fits, one 64-bit atomic per (CTA, id) at the end.
constexpr uint32_t kHistSmemIds = 11264;  // 44 KB of u32 counters
template <bool PRIV>
__global__ void __launch_bounds__(256) token_hist_kernel(const int32_t *__restrict__ tokens, size_t in_stride,
                                                         const int32_t *__restrict__ lens, size_t n_rec, uint32_t n_ids,
                                                         unsigned long long *__restrict__ counts, int *bad) {
    __shared__ uint32_t s_cnt[PRIV ? kHistSmemIds : 1];
    if (PRIV) {
        for (uint32_t i = threadIdx.x; i < n_ids; i += blockDim.x) s_cnt[i] = 0;
        __syncthreads();
    }
    for (size_t r = blockIdx.x; r < n_rec; r += gridDim.x) {
        const int32_t *tok = tokens + r * in_stride;
        const uint32_t n = (uint32_t)min((long long)max(lens[r], 0), (long long)in_stride);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const uint32_t t = (uint32_t)tok[i];
            if (t >= n_ids) { atomicExch(bad, 1); continue; }
            if (PRIV) atomicAdd(&s_cnt[t], 1u);
            else atomicAdd(&counts[t], 1ull);
        }
    }
    if (PRIV) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n_ids; i += blockDim.x)
            if (s_cnt[i]) atomicAdd(&counts[i], (unsigned long long)s_cnt[i]);
    }
}
// ------------------------------------------------------------------ compact (CSR) token output
// The encoder writes int32 rows of out_stride slots; what a host consumer needs is the tokens themselves:
// 2-byte ids (< 65 536 by construction), rows back to back.  One CTA scans the lengths, all CTAs copy.
__global__ void __launch_bounds__(1024) csr_offsets_kernel(const int32_t *__restrict__ len, size_t n_rec, size_t in_stride,
                                                           unsigned long long *__restrict__ off, unsigned long long base) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_base;
    if (threadIdx.x == 0) s_base = base;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (size_t c0 = 0; c0 < n_rec; c0 += 1024) {
        const size_t r = c0 + threadIdx.x;
        unsigned long long x = 0;
        if (r < n_rec) x = (unsigned long long)min((size_t)max(len[r], 0), in_stride);  // tokens beyond the stride were only counted
        unsigned long long incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned long long before = s_base;
        for (int q = 0; q < warp; q++) before += s_warp[q];
        if (r < n_rec) off[r] = before + incl - x;
        __syncthreads();
        if (threadIdx.x == 1023) s_base = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[n_rec] = s_base;
}

__global__ void __launch_bounds__(256) csr_copy_kernel(const int32_t *__restrict__ tok, size_t in_stride, const int32_t *__restrict__ len,
                                                       size_t n_rec, const unsigned long long *__restrict__ off, uint16_t *__restrict__ out) {
    for (size_t r = blockIdx.x; r < n_rec; r += gridDim.x) {
        const size_t n = min((size_t)max(len[r], 0), in_stride);
        const int32_t *src = tok + r * in_stride;
        uint16_t *dst = out + off[r];
        for (size_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = (uint16_t)src[i];
    }
}

}  // namespace ecgb

using namespace ecgb;

extern "C" int ecgb_decode_symbols(const ecgb_vocab *v, const int32_t *d_tokens, size_t n_rec, size_t in_stride,
                                   const int32_t *d_len, uint8_t *d_sym, size_t sym_stride, int32_t *d_sym_len,
                                   void *stream) {
    ECGB_REQUIRE(v, "vocab is NULL");
    if (n_rec == 0) return ECGB_OK;
    ECGB_REQUIRE(d_tokens && d_len && d_sym && d_sym_len, "NULL buffer");
    const VocabView *vv = ecgb_vocab_view(v);
    ECGB_REQUIRE(vv->d_dec_off != nullptr, "vocabulary has no decode table (token ids too large)");
    int device = ecgb_vocab_device(v);
    DeviceGuard g(device);
    cudaStream_t st = as_stream(stream);
    AsyncBuf<int> d_bad;  // released on every exit path
    ECGB_CUDA(d_bad.alloc(1, st));
    ECGB_CUDA(cudaMemsetAsync(d_bad.p, 0, sizeof(int), st));
    decode_kernel<false, uint8_t><<<(unsigned)n_rec, 256, 0, st>>>(d_tokens, in_stride, d_len, d_sym, sym_stride, d_sym_len,
                                                                   vv->d_dec_sym, vv->d_dec_off, vv->dec_ids, d_bad, nullptr);
    ECGB_CUDA(cudaGetLastError());
    int bad = 0;
    ECGB_CUDA(cudaMemcpyAsync(&bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    ECGB_CUDA(cudaStreamSynchronize(st));
    if (bad) return fail(ECGB_EINVAL, "a token id is not in the vocabulary (decode_text would raise KeyError)");
    return ECGB_OK;
}

extern "C" int ecgb_expand_attention(const ecgb_vocab *v, const int32_t *d_tokens, const float *d_attn, size_t n_rec,
                                     size_t in_stride, const int32_t *d_len, float *d_out, size_t out_stride,
                                     int32_t *d_out_len, void *stream) {
    ECGB_REQUIRE(v, "vocab is NULL");
    if (n_rec == 0) return ECGB_OK;
    ECGB_REQUIRE(d_tokens && d_attn && d_len && d_out && d_out_len, "NULL buffer");
    const VocabView *vv = ecgb_vocab_view(v);
    ECGB_REQUIRE(vv->d_dec_off != nullptr, "vocabulary has no decode table (token ids too large)");
    DeviceGuard g(ecgb_vocab_device(v));
    cudaStream_t st = as_stream(stream);
    AsyncBuf<int> d_bad;  // released on every exit path
    ECGB_CUDA(d_bad.alloc(1, st));
    ECGB_CUDA(cudaMemsetAsync(d_bad.p, 0, sizeof(int), st));
    decode_kernel<true, float><<<(unsigned)n_rec, 256, 0, st>>>(d_tokens, in_stride, d_len, d_out, out_stride, d_out_len,
                                                                vv->d_dec_sym, vv->d_dec_off, vv->dec_ids, d_bad, d_attn);
    ECGB_CUDA(cudaGetLastError());
    int bad = 0;
    ECGB_CUDA(cudaMemcpyAsync(&bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    ECGB_CUDA(cudaStreamSynchronize(st));
    if (bad) return fail(ECGB_EINVAL, "a token id is not in the vocabulary (expand_attention would raise KeyError)");
    return ECGB_OK;
}

extern "C" int ecgb_token_histogram(const int32_t *d_tokens, size_t in_stride, const int32_t *d_len, size_t n_rec,
                                    uint32_t n_ids, unsigned long long *d_counts, int device, void *stream) {
    if (n_rec == 0) return ECGB_OK;
    ECGB_REQUIRE(d_tokens && d_len && d_counts, "NULL buffer");
    ECGB_REQUIRE(n_ids > 0, "n_ids is 0");
    DeviceGuard g(device);
    cudaStream_t st = as_stream(stream);
    AsyncBuf<int> d_bad;  // released on every exit path
    ECGB_CUDA(d_bad.alloc(1, st));
    ECGB_CUDA(cudaMemsetAsync(d_bad.p, 0, sizeof(int), st));
    int sms = 0;
    ECGB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const bool priv = n_ids <= kHistSmemIds;
    const unsigned grid = (unsigned)std::min<size_t>(n_rec, (size_t)sms * 4);
    if (priv) token_hist_kernel<true><<<grid, 256, 0, st>>>(d_tokens, in_stride, d_len, n_rec, n_ids, d_counts, d_bad);
    else token_hist_kernel<false><<<grid, 256, 0, st>>>(d_tokens, in_stride, d_len, n_rec, n_ids, d_counts, d_bad);
    ECGB_CUDA(cudaGetLastError());
    int bad = 0;
    ECGB_CUDA(cudaMemcpyAsync(&bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    ECGB_CUDA(cudaStreamSynchronize(st));
    if (bad) return fail(ECGB_EINVAL, "a token id is outside [0, n_ids)");
    return ECGB_OK;
}

extern "C" int ecgb_dequantize(double p1, double p99, const uint8_t *d_sym, size_t n, double *d_out, int device,
                               void *stream) {
    if (n == 0) return ECGB_OK;
    ECGB_REQUIRE(d_sym && d_out, "NULL buffer");
    int rc = check_device(device);
    if (rc) return rc;
    DeviceGuard g(device);
    // tokenizer_utils.py:23-24,27: min = p1 - 0.5, max = p99 + 0.5, span = max - min
    volatile double mn = p1 - 0.5, mx = p99 + 0.5;
    volatile double span = mx - mn;
    const int grid = (int)std::min<size_t>((size_t)sm_count(device) * 8, (n + 255) / 256);
    dequantize_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_sym, n, d_out, mn, span);
    ECGB_CUDA(cudaGetLastError());
    return ECGB_OK;
}

extern "C" int ecgb_pack_training(const int32_t *d_tokens, size_t in_stride, const int32_t *d_len, size_t n_rec,
                                  const int64_t *d_lut, uint32_t lut_size, const int64_t *d_text,
                                  const uint64_t *d_text_off, const int32_t *d_q_len, const ecgb_pack_cfg *cfg,
                                  int64_t *d_input_ids, float *d_attn_mask, int64_t *d_labels, int64_t *d_position_ids,
                                  int32_t *d_status, int device, void *stream) {
    if (n_rec == 0) return ECGB_OK;
    ECGB_REQUIRE(cfg, "cfg is NULL");
    ECGB_REQUIRE(d_tokens && d_len && d_lut && d_text_off && d_q_len && d_input_ids && d_attn_mask && d_labels &&
                     d_position_ids && d_status, "NULL buffer");
    int rc = check_device(device);
    if (rc) return rc;
    DeviceGuard g(device);
    PackCfg c{cfg->pad_id, cfg->bos_id, cfg->eos_id, cfg->sig_start_id, cfg->sig_end_id, cfg->pad_to_max};
    pack_kernel<<<(unsigned)n_rec, 256, 0, as_stream(stream)>>>(
        d_tokens, in_stride, d_len, reinterpret_cast<const long long *>(d_lut), lut_size,
        reinterpret_cast<const long long *>(d_text), reinterpret_cast<const unsigned long long *>(d_text_off), d_q_len, c,
        reinterpret_cast<long long *>(d_input_ids), d_attn_mask, reinterpret_cast<long long *>(d_labels),
        reinterpret_cast<long long *>(d_position_ids), d_status);
    ECGB_CUDA(cudaGetLastError());
    return ECGB_OK;
}

// Compact copy of the encoder's output: rows of 2-byte token ids back to back (CSR).  d_off[n_rec + 1] receives
// the row offsets in tokens, counted from `base` (d_off[n_rec] = base + total); row r goes to out[d_off[r] ...], so
// `out` needs room for base + sum(min(len, in_stride)) tokens.  `out` may be PINNED HOST memory (mapped into the
// device's address space, as cudaHostAlloc / torch pin_memory() memory is): the copy kernel then stores the tokens
// straight over PCIe and no host-side size hand-over is needed.
extern "C" int ecgb_tokens_csr(const int32_t *d_tokens, size_t in_stride, const int32_t *d_len, size_t n_rec, uint16_t *d_out,
                               uint64_t *d_off, uint64_t base, int device, void *stream) {
    ECGB_REQUIRE(d_off, "d_off is NULL");
    int rc = check_device(device);
    if (rc) return rc;
    DeviceGuard g(device);
    cudaStream_t st = as_stream(stream);
    ECGB_REQUIRE(n_rec == 0 || (d_tokens && d_len && d_out), "NULL buffer");
    csr_offsets_kernel<<<1, 1024, 0, st>>>(d_len, n_rec, in_stride, reinterpret_cast<unsigned long long *>(d_off), (unsigned long long)base);
    if (n_rec) {
        const unsigned grid = (unsigned)std::min<size_t>(n_rec, (size_t)sm_count(device) * 8);
        csr_copy_kernel<<<grid, 256, 0, st>>>(d_tokens, in_stride, d_len, n_rec, reinterpret_cast<const unsigned long long *>(d_off), d_out);
    }
    ECGB_CUDA(cudaGetLastError());
    return ECGB_OK;
}
