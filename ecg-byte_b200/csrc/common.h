// Shared host-side helpers of libecgbyte.so (status codes, error string, CUDA checks).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "ecgbyte.h"

namespace ecgb {

// thread-local message returned by ecgb_last_error()
char *err_buf();
int fail(int status, const char *fmt, ...);

// Stream-ordered scratch allocation that is returned to the pool on EVERY exit path of the function that owns it
// (the ECGB_CUDA / ECGB_REQUIRE macros return early on failure).
template <class T>
struct AsyncBuf {
    T *p = nullptr;
    cudaStream_t st = nullptr;
    AsyncBuf() = default;
    AsyncBuf(const AsyncBuf &) = delete;
    AsyncBuf &operator=(const AsyncBuf &) = delete;
    cudaError_t alloc(size_t count, cudaStream_t stream) {
        st = stream;
        return cudaMallocAsync(reinterpret_cast<void **>(&p), (count ? count : 1) * sizeof(T), stream);
    }
    ~AsyncBuf() {
        if (p) cudaFreeAsync(p, st);
    }
    operator T *() const { return p; }
};

#define ECGB_CUDA(call)                                                                      \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return ::ecgb::fail(e_ == cudaErrorMemoryAllocation ? ECGB_ENOMEM : ECGB_ECUDA,  \
                                "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),      \
                                __FILE__, __LINE__);                                         \
    } while (0)

#define ECGB_REQUIRE(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) return ::ecgb::fail(ECGB_EINVAL, __VA_ARGS__); \
    } while (0)

// Makes `device` current for the scope (restores the previous one).
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int check_device(int device);  // ECGB_OK or ECGB_ENODEVICE / ECGB_EINVAL
int sm_count(int device);

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- internal views shared between translation units ----

constexpr int kNumSymbols = 26;    // len(ALPHABET), tokenizer_utils.py:12
constexpr int kNumThresholds = 25;
constexpr int kCells = 32;         // pre-classification cells of the quantiser (26 used)

// Device-resident quantiser tables.  A sample s is mapped to a cell by monotone fp32
// arithmetic  cell = min(uint((float(s) - lo) * scale), 25); the cells are half a bin out of
// phase with the symbol bins, so cell k holds exactly threshold t_{k+1} and
// symbol = cell + (s >= thr_cell[cell])  -- one 4-byte shared-memory load per sample.
struct QuantTables {
    float lo, scale;
    const void *d_cell_thr;     // ThrT[kCells]  (float for f32/i16, double for f64); [25..] = NaN
    const void *d_thr;          // ThrT[kNumThresholds + 2] with sentinels (generic path)
    int exact_cells;            // 1: the cell table is usable; 0: fall back to d_thr search
};

// device image of the pair table (trie_host.h); d_ent == nullptr when the vocabulary has none
struct PairView {
    const uint32_t *d_ent;
    const uint16_t *d_tok;
    const uint8_t *d_cls;   // byte -> class, SE for bytes without one
    uint32_t n_ent, root_base, W, SM, SE;
};

struct VocabView {
    const uint2 *d_nodes;   // compact nodes: x = child mask (bit c = class c), y = base << 16 | (token + 1)
    uint32_t n_nodes;
    uint32_t smem_nodes;    // leading nodes staged in shared memory
    const uint8_t *d_cls;   // byte -> class (0..30) or 31 = no child anywhere
    const uint32_t *d_wide; // wide nodes (10 words each) when !compact
    int compact;
    int ecg_alphabet;       // class('a'+k) == k for k < 26
    uint32_t max_token_len;
    // decode tables (tokenizer_utils.py:75-77): token id t expands to d_dec_sym[d_dec_off[t] .. d_dec_off[t+1])
    const uint8_t *d_dec_sym;
    const uint32_t *d_dec_off;
    uint32_t dec_ids;       // ids 0 .. dec_ids-1 are covered by d_dec_off
    PairView pair;
};

}  // namespace ecgb

struct ecgb_quantizer {
    int device;
    ecgb_dtype dtype;
    double p1, p99, i16_scale, lo, den;
    double thr[ecgb::kNumThresholds];  // thresholds in the sample domain, as doubles
    ecgb::QuantTables tab;
    void *d_block;  // one allocation holding all tables
};

namespace ecgb {
// encode2.cu: the pair-table walker; ECGB_EUNSUPPORTED = take the bitmap-trie kernel instead
int launch_encode2(int dt, const VocabView *vv, const QuantTables *qt, int exact_cells, const void *d_in, size_t n_total,
                   size_t n_rec, size_t rec_len, const uint64_t *d_offsets, int32_t *d_tokens, size_t out_stride,
                   int32_t *d_len, int device, cudaStream_t st);
}  // namespace ecgb

const ecgb::VocabView *ecgb_vocab_view(const ecgb_vocab *v);
// one long string, parallel over positions (encode_long.cu); compact vocabularies, n < 2^32
int ecgb_encode_long_device(const ecgb_vocab *v, const uint8_t *d_text, size_t n, uint32_t *d_out, size_t cap,
                            unsigned long long *h_count, cudaStream_t st);
int ecgb_vocab_device(const ecgb_vocab *v);
