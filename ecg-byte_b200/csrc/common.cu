// Error reporting, device queries and small host utilities of libecgbyte.so.
#include "common.h"

#include <cstring>
#include <vector>

namespace ecgb {

char *err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int status, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return status;
}

int check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(ECGB_ENODEVICE, "no usable CUDA device (%s); libecgbyte has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n) return fail(ECGB_EINVAL, "device %d out of range [0, %d)", device, n);
    return ECGB_OK;
}

int sm_count(int device) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
    return n;
}

}  // namespace ecgb

using namespace ecgb;

extern "C" const char *ecgb_last_error(void) { return err_buf(); }
extern "C" int ecgb_version(void) { return 100; }

extern "C" int ecgb_device_count(int *n_out) {
    ECGB_REQUIRE(n_out, "n_out is NULL");
    *n_out = 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(ECGB_ENODEVICE, "no usable CUDA device (%s)", e != cudaSuccess ? cudaGetErrorString(e) : "count 0");
    }
    *n_out = n;
    return ECGB_OK;
}

// lib.rs:101-110: vocab_tokens[new] = expand(left) ++ expand(right)
extern "C" int ecgb_expand_merges(const uint32_t *h_pairs, uint32_t n_merges, uint32_t *h_seq, uint64_t seq_cap,
                                  uint64_t *h_seq_off) {
    ECGB_REQUIRE(h_seq_off && (h_pairs || n_merges == 0), "NULL argument");
    std::vector<uint64_t> len((size_t)n_merges + 256, 1);
    h_seq_off[0] = 0;
    for (uint32_t i = 0; i < n_merges; i++) {
        uint32_t l = h_pairs[2 * i], r = h_pairs[2 * i + 1];
        ECGB_REQUIRE(l < 256 + i && r < 256 + i, "merge %u refers to a token that does not exist yet", i);
        len[256 + i] = len[l] + len[r];
        h_seq_off[i + 1] = h_seq_off[i] + len[256 + i];
    }
    if (h_seq_off[n_merges] > seq_cap || (!h_seq && h_seq_off[n_merges] > 0))
        return fail(ECGB_ECAPACITY, "sequence buffer too small: need %llu", (unsigned long long)h_seq_off[n_merges]);
    for (uint32_t i = 0; i < n_merges; i++) {
        uint32_t *dst = h_seq + h_seq_off[i];
        const uint32_t side[2] = {h_pairs[2 * i], h_pairs[2 * i + 1]};
        for (int s = 0; s < 2; s++) {
            uint32_t t = side[s];
            if (t < 256) {
                *dst++ = t;
            } else {
                uint64_t o = h_seq_off[t - 256], e = h_seq_off[t - 256 + 1];
                std::memcpy(dst, h_seq + o, (size_t)(e - o) * sizeof(uint32_t));
                dst += e - o;
            }
        }
    }
    return ECGB_OK;
}
