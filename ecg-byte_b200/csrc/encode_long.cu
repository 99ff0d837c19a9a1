// E2 for ONE long string (rust_bpe.encode_text on a whole record or corpus string,
// lib.rs:149-193).  The walker-per-record kernel would run such a call on a single thread;
// here the string itself is the parallel axis:
//   A  every position p computes its own longest match (len[p], tok[p])  -- embarrassingly
//      parallel, ~36 trie steps per position on ECG text;
//   B  per chunk of 2048 positions, backwards: exit[p] = first position of the token chain
//      starting at p that lies beyond the chunk (chains from different entries merge fast,
//      but no such assumption is made -- every position gets its exact exit);
//   C  one thread hops chunk to chunk: entry[k+1] = exit[entry[k]];
//   D  per chunk: follow the chain from its entry, count, exclusive-scan the counts, write.
// The result is the greedy longest-match tokenisation, identical to the sequential walk.
#include <algorithm>

#include "common.h"

namespace ecgb {

constexpr int kChunk = 2048;
constexpr uint32_t kNone = 0xFFFFFFFFu;

__global__ void __launch_bounds__(256) long_match_kernel(const uint8_t *__restrict__ text, size_t n,
                                                         const uint2 *__restrict__ nodes, const uint8_t *__restrict__ cls,
                                                         uint16_t *__restrict__ len, uint16_t *__restrict__ tok) {
    __shared__ uint8_t s_cls[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_cls[i] = cls[i];
    __syncthreads();
    const uint2 root = __ldg(nodes);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        uint32_t mask = root.x, base = root.y >> 16, best_len = 0, best_tok = 0;
        for (size_t j = p; j < n; j++) {
            const uint32_t c = s_cls[text[j]];
            if (c >= 31u || !((mask >> c) & 1u)) break;
            const uint2 nd = __ldg(nodes + base + __popc(mask & ((1u << c) - 1u)));
            mask = nd.x;
            base = nd.y >> 16;
            const uint32_t t = nd.y & 0xFFFFu;
            if (t) { best_len = (uint32_t)(j - p) + 1u; best_tok = t - 1u; }
            if (best_len == 0xFFFFu) break;  // lengths are stored in 16 bits (vocab_create guarantees max_token_len fits)
        }
        if (best_len == 0) { best_len = 1; best_tok = text[p]; }  // a byte that occurs in no merge: its own token
        len[p] = (uint16_t)best_len;
        tok[p] = (uint16_t)best_tok;
    }
}

// exit[p] (relative to nothing: absolute position, 32-bit) for every p of chunk k, backwards
__global__ void __launch_bounds__(128) long_exit_kernel(const uint16_t *__restrict__ len, size_t n, uint32_t *__restrict__ exitp,
                                                        size_t n_chunks) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_chunks) return;
    const size_t lo = k * kChunk, hi = (lo + (size_t)kChunk < n) ? lo + (size_t)kChunk : n;
    for (size_t p = hi; p-- > lo;) {
        const size_t e = p + len[p];
        exitp[p] = e >= hi ? (uint32_t)e : exitp[e];
    }
}

// entry[k] = first token start inside chunk k (kNone if a long token skips the chunk)
__global__ void long_chain_kernel(const uint32_t *__restrict__ exitp, size_t n, uint32_t *__restrict__ entry, size_t n_chunks) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (size_t k = 0; k < n_chunks; k++) entry[k] = kNone;
    size_t pos = 0;
    while (pos < n) {
        entry[pos / kChunk] = (uint32_t)pos;
        pos = exitp[pos];
    }
}

// pass 0: count the tokens of each chunk; pass 1: write them at the scanned offsets
__global__ void __launch_bounds__(128) long_emit_kernel(const uint16_t *__restrict__ len, const uint16_t *__restrict__ tok, size_t n,
                                                        const uint32_t *__restrict__ entry, size_t n_chunks,
                                                        unsigned long long *__restrict__ counts, uint32_t *__restrict__ out,
                                                        size_t cap, int pass) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_chunks) return;
    const size_t hi = ((k + 1) * (size_t)kChunk < n) ? (k + 1) * (size_t)kChunk : n;
    size_t pos = entry[k];
    if (pass == 0) {
        unsigned long long c = 0;
        if (pos != kNone)
            for (; pos < hi; pos += len[pos]) c++;
        counts[k] = c;
    } else if (pos != kNone) {
        unsigned long long o = counts[k];
        for (; pos < hi; pos += len[pos], o++)
            if (o < cap) out[o] = tok[pos];
    }
}

// in-place exclusive scan of counts[0..m) by one block; total in counts[m]
__global__ void __launch_bounds__(1024) long_scan_kernel(unsigned long long *counts, size_t m) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (size_t base = 0; base < m; base += blockDim.x) {
        const size_t i = base + threadIdx.x;
        const unsigned long long v = i < m ? counts[i] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned long long woff = 0, total = 0;
        for (int w = 0; w < 32; w++) { if (w < warp) woff += s_warp[w]; total += s_warp[w]; }
        if (i < m) counts[i] = s_carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[m] = s_carry;
}

}  // namespace ecgb

using namespace ecgb;

// Encodes d_text[0..n) (device) into d_out (device, capacity cap tokens); *h_count = true token count.
int ecgb_encode_long_device(const ecgb_vocab *v, const uint8_t *d_text, size_t n, uint32_t *d_out, size_t cap,
                            unsigned long long *h_count, cudaStream_t st) {
    const VocabView *vv = ecgb_vocab_view(v);
    const int device = ecgb_vocab_device(v);
    const int sms = sm_count(device);
    const size_t n_chunks = (n + kChunk - 1) / kChunk;
    // scratch from the stream's pool; released on every exit path
    AsyncBuf<uint16_t> d_len, d_tok;
    AsyncBuf<uint32_t> d_exit, d_entry;
    AsyncBuf<unsigned long long> d_counts;
    ECGB_CUDA(d_len.alloc(n, st));
    ECGB_CUDA(d_tok.alloc(n, st));
    ECGB_CUDA(d_exit.alloc(n, st));
    ECGB_CUDA(d_entry.alloc(n_chunks, st));
    ECGB_CUDA(d_counts.alloc(n_chunks + 1, st));
    const int gridA = (int)std::min<size_t>((size_t)sms * 8, (n + 255) / 256);
    long_match_kernel<<<gridA, 256, 0, st>>>(d_text, n, vv->d_nodes, vv->d_cls, d_len, d_tok);
    const int gridC = (int)((n_chunks + 127) / 128);
    long_exit_kernel<<<gridC, 128, 0, st>>>(d_len, n, d_exit, n_chunks);
    long_chain_kernel<<<1, 32, 0, st>>>(d_exit, n, d_entry, n_chunks);
    long_emit_kernel<<<gridC, 128, 0, st>>>(d_len, d_tok, n, d_entry, n_chunks, d_counts, d_out, cap, 0);
    long_scan_kernel<<<1, 1024, 0, st>>>(d_counts, n_chunks);
    long_emit_kernel<<<gridC, 128, 0, st>>>(d_len, d_tok, n, d_entry, n_chunks, d_counts, d_out, cap, 1);
    ECGB_CUDA(cudaGetLastError());
    ECGB_CUDA(cudaMemcpyAsync(h_count, d_counts.p + n_chunks, 8, cudaMemcpyDeviceToHost, st));
    ECGB_CUDA(cudaStreamSynchronize(st));
    return ECGB_OK;
}
