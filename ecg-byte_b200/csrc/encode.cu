// E2: encode_text (reference: rust_bpe/src/lib.rs:149-193) fused with Q1
// (tokenizer_utils.py:14-19) for sm_100a.
//
// Semantics: greedy longest match over the trie -- at each token start walk as far
// as symbols match, remember the longest terminal seen, emit it and restart right
// after it (lib.rs:163-190).  This is NOT rank-ordered BPE merging.
//
// Mapping (one walker = one thread = one record, 1 CTA of up to 768 walkers per SM, 85 registers each):
//   * records are independent and 100k-1M of them are in flight, so the record axis alone fills
//     the chip: no speculation, no redundant trie steps, no inter-thread synchronisation after the tables are staged;
//   * the trie (8-byte bitmap nodes, 104 KB for 5,000 merges) is staged once per CTA in shared memory (nodes
//     that do not fit stay in L2; tables that spill go to the pair-table walker of encode2.cu instead);
//   * a warp alternates between two CONVERGENT phases.  Refill: every lane with room loads 16 samples of its own
//     record (4 x LDG.128), classifies them (K1's arithmetic, quant_device.cuh) and appends 16 shift codes to its
//     private 128-symbol ring in shared memory; rings are transposed per warp (word w of lane l at (w * 32 + l) * 4),
//     so ring traffic is conflict-free wherever each cursor is.  Walk: K = REDUX.MIN(symbols left in the rings)
//     branch-free trie steps -- one LDS.U8 (symbol, loaded one step ahead), one LDS.64 (node), a sign test and a
//     POPC per step, predicated token store; the end of a record is a sentinel symbol that has no edge;
//   * a failed walk restarts right after the emitted token at a ring position that is still there (the ring never
//     drops symbols after the last terminal); the rare walk longer than the ring finishes symbol by symbol from
//     global memory.
// HBM traffic is the algorithmic minimum: samples once, tokens once (ncu: 270.7 KB per record against 258.3 KB).
// Measured history and the limits of this design: DESIGN.md section 4, K2.
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

constexpr uint32_t kNoClass = 31u;
// The rings hold SHIFT CODES, 31 - class: `mask << code` moves the symbol's edge bit to bit 31 (the
// match test is a sign test) and the lower classes above bit 31 - class (one more shift and a POPC give
// the child rank).  The sentinel class 31 is code 0; bit 31 of a child bitmap is never set.
constexpr uint32_t kNoCode = 31u - kNoClass;

// ---- 16-sample groups: the unit a walker loads, quantises and appends to its ring ----
template <int DT> struct ElemOf { using T = typename SampleTraits<DT>::In; };
template <> struct ElemOf<ECGB_U8> { using T = uint8_t; };
template <int DT> struct ThrOf { using T = typename SampleTraits<DT>::Thr; };
template <> struct ThrOf<ECGB_U8> { using T = float; };

constexpr int kGroup = 16;
#ifndef ECGB_ENC1_MINLANES
#define ECGB_ENC1_MINLANES 20
#endif
#ifndef ECGB_ENC1_NEEDY
#define ECGB_ENC1_NEEDY 48
#endif
constexpr int kMinRefillLanes = ECGB_ENC1_MINLANES;  // refill rounds wanted by fewer lanes wait ...
constexpr int kNeedySymbols = ECGB_ENC1_NEEDY;       // ... unless a lane has fewer symbols than this ahead of its cursor
constexpr int kMaxThreads = 768;  // walkers per CTA: 85 registers each, 24 warps per SM

// 16 samples at global index g (g % 16 == 0) -> 16 symbol classes, one per byte.
// Positions at or beyond `valid` (record end) become the sentinel class 31, which has no
// edge anywhere in the trie and therefore ends every walk.
template <int DT, bool CELLS>
__device__ __forceinline__ uint4 fetch16(const void *base, size_t g, size_t n_total, int valid, const void *qsmem,
                                         const void *thr_smem, const uint8_t *s_cls, float lo, float scale) {
    using T = typename ElemOf<DT>::T;
    const T *p = static_cast<const T *>(base) + g;
    constexpr int NV = sizeof(T);  // 16-byte vectors per group
    uint4 raw[NV];
    if (g + kGroup <= n_total) {
#pragma unroll
        for (int j = 0; j < NV; j++) raw[j] = __ldg(reinterpret_cast<const uint4 *>(p) + j);
    } else {  // ragged end of the buffer: element-wise, zero filled
        T tmp[kGroup];
#pragma unroll
        for (int k = 0; k < kGroup; k++) tmp[k] = (g + k < n_total) ? p[k] : T(0);
        memcpy(raw, tmp, sizeof(raw));
    }
    const T *e = reinterpret_cast<const T *>(raw);
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < kGroup; k++) {
        uint32_t q;
        if constexpr (DT == ECGB_U8) {
            q = s_cls[e[k]];
        } else {
            using Thr = typename SampleTraits<DT>::Thr;
            float sf;
            Thr s = to_thr(e[k], &sf);
            q = CELLS ? classify<Thr>(s, sf, lo, scale, static_cast<const QuantSmem<Thr> *>(qsmem))
                      : classify_search<Thr>(s, static_cast<const Thr *>(thr_smem));
        }
        w[k >> 2] |= q << ((k & 3) * 8);
    }
    if (valid < kGroup) {  // the record ends inside (or before) this group
#pragma unroll
        for (int k = 0; k < kGroup; k++)
            if (k >= valid) w[k >> 2] |= kNoClass << ((k & 3) * 8);  // classes are < 32: OR gives 31
    }
    // classes -> shift codes, four at a time (every byte is <= 31: no borrow between bytes)
#pragma unroll
    for (int j = 0; j < 4; j++) w[j] = 0x1F1F1F1Fu - w[j];
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Symbol rings, one per walker, R symbols each, laid out TRANSPOSED inside a warp's
// region: 4-byte word w of lane l lives at (w * 32 + l) * 4, so the bank of every access
// is the lane id -- ring reads and writes are conflict-free wherever each lane's cursor is.
template <int R>
__device__ __forceinline__ uint32_t ring_addr(uint32_t lane_base, int32_t pos) {
    const uint32_t o = (uint32_t)pos & (uint32_t)(R - 1);
    return lane_base + ((o & ~3u) << 5) + (o & 3u);
}

// class of the single symbol at global index g (slow paths only)
template <int DT, bool CELLS>
__device__ __forceinline__ uint32_t class_at(const void *base, size_t g, const void *qsmem, const void *thr_smem,
                                             const uint8_t *s_cls, float lo, float scale) {
    using T = typename ElemOf<DT>::T;
    const T v = static_cast<const T *>(base)[g];
    if constexpr (DT == ECGB_U8) {
        return s_cls[v];
    } else {
        using Thr = typename SampleTraits<DT>::Thr;
        float sf;
        Thr s = to_thr(v, &sf);
        return CELLS ? classify<Thr>(s, sf, lo, scale, static_cast<const QuantSmem<Thr> *>(qsmem))
                     : classify_search<Thr>(s, static_cast<const Thr *>(thr_smem));
    }
}

// One walker (thread) per record.  A warp alternates between two CONVERGENT phases so
// that the 32 walkers never serialise on each other's bookkeeping:
//   refill: every lane with room appends 16-sample groups to its private ring (128-bit
//           loads -> threshold classification -> four conflict-free STS.32 per group);
//           record ends, record switches and the (rare) walks longer than a ring are
//           handled here, outside the hot loop;
//   walk:   K = min over lanes of the symbols left in their rings (one REDUX) trie steps
//           with no votes, no bounds checks and no branches inside: one LDS.U8 symbol +
//           one LDS.64 node per step, token emission predicated.  The end of a record is
//           a sentinel symbol in the ring (class 31 has no edge anywhere), on which a
//           finished walker simply idles until the phase ends.
// All lanes consume ~1 symbol per step, so their rings drain in lockstep and nearly all
// lanes take part in every refill.
template <int DT, bool CELLS, bool ALL_SMEM, int R>
__global__ void __launch_bounds__(kMaxThreads, 1) encode_kernel(EncArgs a) {
    using Thr = typename ThrOf<DT>::T;
    extern __shared__ __align__(16) uint8_t smem[];
    uint2 *s_nodes = reinterpret_cast<uint2 *>(smem);
    const size_t nodes_bytes = ((size_t)a.smem_nodes * 8 + 15) & ~(size_t)15;
    uint8_t *s_aux = smem + nodes_bytes;
    // aux region: quantiser tables (sample dtypes) or the byte->class table (text)
    QuantSmem<Thr> *qs = reinterpret_cast<QuantSmem<Thr> *>(s_aux);
    Thr *s_thr = reinterpret_cast<Thr *>(s_aux + sizeof(QuantSmem<Thr>));
    uint8_t *s_cls = s_aux;
    constexpr size_t kAux = DT == ECGB_U8 ? 256 : sizeof(QuantSmem<Thr>) + sizeof(Thr) * 32;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(s_aux + kAux) + (threadIdx.x >> 5) * (32u * R) +
                          (threadIdx.x & 31u) * 4u;
    const uint32_t nodes_sa = (uint32_t)__cvta_generic_to_shared(s_nodes);

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
    const uint32_t stride = (uint32_t)a.out_stride;
    // explicit offsets: the buffer ends where the last record ends
    const size_t n_total = a.offsets ? (size_t)a.offsets[a.n_rec] : a.n_total;
    constexpr unsigned FULL = 0xffffffffu;
    const uint2 root = s_nodes[0];
    // breadth-first numbering: only the root has its first child at index 1, so
    // "a walk is in progress" == (base != root_base)
    const uint32_t root_base = root.y >> 16;

    auto node_at = [&](uint32_t idx) -> uint2 {
        if (ALL_SMEM || idx < S) return lds_u64(nodes_sa + idx * 8u);
        return __ldg(a.nodes + idx);
    };

    // contiguous, even split of the records over the CTAs
    const size_t r_lo = (size_t)(((unsigned __int128)a.n_rec * blockIdx.x) / gridDim.x);
    const size_t r_hi = (size_t)(((unsigned __int128)a.n_rec * (blockIdx.x + 1)) / gridDim.x);
    size_t r_next = r_lo + threadIdx.x;

    // walker state; positions are relative to `org` (record start rounded down to a group)
    bool active = false, done = false;
    size_t org = 0, r_cur = 0;
    int32_t end32 = 0;  // record end
    int32_t pos32 = 0;  // next symbol to read
    int32_t mpos = 0;   // end of the longest terminal of the walk in progress (== pos32 at the root)
    int32_t hi32 = 0;   // ring holds [max(hi32 - R, 0), hi32) and never drops symbols >= mpos
    uint32_t mask = root.x, base = root_base, mid = 0, cnt = 0;
    int32_t *outp = nullptr;
    sts_u8(ring, kNoCode);  // idle lanes sit on a sentinel

    for (;;) {
        // ------------------------------------------------ phase boundary (convergent)
        // (1) walkers parked on a class-31 symbol at the root: end of record, or a byte that
        //     occurs in no merge and is its own token (text only; lib.rs:155-157)
        while (active && base == root_base && pos32 < hi32 && lds_u8(ring_addr<R>(ring, pos32)) == kNoCode) {
            if (pos32 >= end32) {
                a.lens[r_cur] = (int32_t)cnt;
                active = false;
            } else {
                const uint32_t byte = static_cast<const uint8_t *>(a.in)[org + (size_t)pos32];
                if (cnt < stride) outp[cnt] = (int32_t)byte;
                cnt++;
                pos32++;
                mpos = pos32;
            }
        }
        // (2) next record
        if (!active && !done) {
            if (r_next < r_hi) {
                r_cur = r_next;
                r_next += blockDim.x;
                const size_t rs = a.offsets ? (size_t)a.offsets[r_cur] : r_cur * a.rec_len;
                const size_t re = a.offsets ? (size_t)a.offsets[r_cur + 1] : rs + a.rec_len;
                org = rs & ~(size_t)(kGroup - 1);
                end32 = (int32_t)(re - org);
                pos32 = mpos = (int32_t)(rs - org);
                hi32 = 0;
                mask = root.x;
                base = root_base;
                mid = cnt = 0;
                outp = a.tokens + r_cur * a.out_stride;
                active = true;
            } else {
                done = true;
                pos32 = mpos = hi32 = end32 = 0;
                mask = root.x;
                base = root_base;
                sts_u8(ring, kNoCode);
            }
        }
        if (__all_sync(FULL, done)) break;
        // (3) refill: append groups while the ring keeps everything from mpos on; the group that
        //     holds the end-of-record sentinel is part of the stream
        bool again;
        do {
#pragma unroll 1
            for (int g = 0; g < R / kGroup; g++) {
                const bool want = active && hi32 <= end32 && hi32 + kGroup - (mpos & ~(kGroup - 1)) <= R;
                // a round costs the same for one lane as for 32: rounds that only a few lanes want are put off until
                // they fill up, unless one of those lanes is about to run dry (see encode2.cu)
                const unsigned wm = __ballot_sync(FULL, want);
                if (!wm) break;
                if (__popc(wm) < kMinRefillLanes && !__any_sync(FULL, want && hi32 - pos32 < kNeedySymbols)) break;
                if (want) {
                    const uint4 sy = fetch16<DT, CELLS>(a.in, org + (size_t)hi32, n_total, end32 - hi32, qs, s_thr,
                                                        s_cls, qlo, qscale);
                    const uint32_t wa = ring_addr<R>(ring, hi32);
                    sts_u32(wa, sy.x);
                    sts_u32(wa + 128u, sy.y);
                    sts_u32(wa + 256u, sy.z);
                    sts_u32(wa + 384u, sy.w);
                    hi32 += kGroup;
                }
            }
            // (4) a walk longer than the ring (no terminal for ~R symbols): finish that one walk
            //     symbol by symbol from global memory, emit, and restart with an empty ring
            const bool stuck = active && pos32 >= hi32 && hi32 <= end32;
            if (stuck) {
                for (;;) {
                    const uint32_t c = pos32 < end32
                                           ? class_at<DT, CELLS>(a.in, org + (size_t)pos32, qs, s_thr, s_cls, qlo, qscale)
                                           : kNoClass;
                    const uint32_t bit = 1u << c;
                    if (!(mask & bit)) break;
                    const uint2 nd = node_at(base + __popc(mask & (bit - 1u)));
                    pos32++;
                    if (nd.y & 0xFFFFu) { mpos = pos32; mid = (nd.y & 0xFFFFu) - 1u; }
                    mask = nd.x;
                    base = nd.y >> 16;
                }
                if (cnt < stride) outp[cnt] = (int32_t)mid;
                cnt++;
                const int32_t shift = mpos & ~(kGroup - 1);
                org += (size_t)shift;
                end32 -= shift;
                mpos -= shift;
                pos32 = mpos;
                hi32 = 0;
                mask = root.x;
                base = root_base;
            }
            again = __any_sync(FULL, stuck);
        } while (again);

        // ------------------------------------------------ walk (convergent trie steps)
        {
            // idle lanes (done) never limit K; a lane that sits on its end-of-record sentinel
            // ends the phase at once.  After the K steps control returns to the phase boundary,
            // which also serves walkers parked on a byte that is its own token.
            uint32_t avail = 0x7fffffffu;
            if (!done) {
                avail = active ? (uint32_t)max(hi32 - pos32, 0) : 0u;
                if (active && base == root_base && pos32 >= end32) avail = 0u;
            }
            uint32_t K = __reduce_min_sync(FULL, avail);
            // The symbol of the next step is loaded one step ahead (ring[pos + 1], speculating on a
            // match) and the symbol at the restart point is remembered when a terminal is passed, so
            // no shared-memory load sits on the step-to-step dependency chain.
            uint32_t c = lds_u8(ring_addr<R>(ring, pos32));  // shift code of the symbol at pos32
            uint32_t cm = lds_u8(ring_addr<R>(ring, mpos));  // ... at mpos
#pragma unroll 2
            for (; K > 0; K--) {
                const int32_t adv = pos32 + 1;
                const uint32_t cspec = lds_u8(ring_addr<R>(ring, adv));
                const uint32_t sh = mask << c;  // edge bit of the symbol at bit 31; the sentinel never has one
                const bool okm = (int32_t)sh < 0;
                // a failed step re-reads the root, which is also the state a new token starts from
                const uint32_t idx = okm ? base + __popc(sh << 1) : 0u;
                const uint2 nd = node_at(idx);
                const bool emit = !okm && base != root_base;  // walk ended: emit the longest terminal
                if (emit && cnt < stride) outp[cnt] = (int32_t)mid;
                cnt += emit ? 1u : 0u;
                const uint32_t tok1 = nd.y & 0xFFFFu;  // token id + 1, 0 = not a token
                if (okm && tok1 != 0u) { mpos = adv; mid = tok1 - 1u; cm = cspec; }
                pos32 = okm ? adv : mpos;  // at the root mpos == pos32: a parked walker stays put
                c = okm ? cspec : cm;
                mask = nd.x;
                base = nd.y >> 16;
            }
        }
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

template <int DT, int R>
static auto pick_kernel(bool cells, bool all_smem) {
    return cells ? (all_smem ? encode_kernel<DT, true, true, R> : encode_kernel<DT, true, false, R>)
                 : (all_smem ? encode_kernel<DT, false, true, R> : encode_kernel<DT, false, false, R>);
}

template <int DT>
static int launch_encode_t(const EncArgs &a, int exact_cells, int device, cudaStream_t st) {
    using Thr = typename ThrOf<DT>::T;
    int sms = sm_count(device);
    int smem_max = 0;
    ECGB_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    const size_t aux = DT == ECGB_U8 ? 256 : sizeof(QuantSmem<Thr>) + sizeof(Thr) * 32;
    EncArgs args = a;
    // one walker per record; CTAs get equal contiguous record ranges
    size_t grid = std::min<size_t>((size_t)sms, (a.n_rec + 31) / 32);
    if (grid < 1) grid = 1;
    size_t per_cta = (a.n_rec + grid - 1) / grid;
    int block = (int)std::min<size_t>(kMaxThreads, ((per_cta + 31) / 32) * 32);
    // 128-symbol rings when the whole trie still fits beside them, else 64-symbol rings
    const size_t nodes_all = (((size_t)a.n_nodes * 8 + 15) & ~(size_t)15);
    int ring = 128;
    if (nodes_all + aux + (size_t)block * 128 + 1024 > (size_t)smem_max) ring = 64;
    const size_t rings = (size_t)block * ring;
    size_t budget = (size_t)smem_max > aux + rings + 1024 ? (size_t)smem_max - aux - rings - 1024 : 0;
    args.smem_nodes = (uint32_t)std::min<size_t>(a.n_nodes, budget / 8);
    if (args.smem_nodes < 1) return fail(ECGB_EUNSUPPORTED, "device shared memory too small for the trie root");
    size_t smem = (((size_t)args.smem_nodes * 8 + 15) & ~(size_t)15) + aux + rings;
    const bool all_smem = args.smem_nodes == a.n_nodes;
    auto kern = ring == 128 ? pick_kernel<DT, 128>(exact_cells != 0, all_smem) : pick_kernel<DT, 64>(exact_cells != 0, all_smem);
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
    ECGB_REQUIRE(((uintptr_t)d_sym & 15) == 0, "d_sym must be 16-byte aligned");
    ECGB_REQUIRE(rec_len < (1ull << 31), "records longer than 2^31 symbols are not supported");
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
    {
        const int rc2 = launch_encode2(ECGB_U8, vv, nullptr, 1, d_sym, n_rec * rec_len, n_rec, rec_len, d_offsets, d_tokens,
                                       out_stride, d_len, device, st);
        if (rc2 != ECGB_EUNSUPPORTED) return rc2;
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
    ECGB_REQUIRE(rec_len < (1ull << 31), "records longer than 2^31 samples are not supported");
    const VocabView *vv = ecgb_vocab_view(v);
    int device = ecgb_vocab_device(v);
    ECGB_REQUIRE(device == q->device, "vocab (device %d) and quantizer (device %d) live on different devices", device, q->device);
    DeviceGuard g(device);
    cudaStream_t st = as_stream(stream);
    if (!vv->compact) {
        // wide vocabularies: quantise to a temporary symbol buffer, then the wide walker
        AsyncBuf<uint8_t> d_sym;  // released on every exit path
        ECGB_CUDA(d_sym.alloc(n_rec * rec_len, st));
        int rc = ecgb_quantize(q, d_in, n_rec * rec_len, d_sym, stream);
        if (rc == ECGB_OK) rc = ecgb_encode_symbols(v, d_sym, n_rec, rec_len, nullptr, d_tokens, out_stride, d_len, stream);
        return rc;
    }
    {
        const int rc2 = launch_encode2((int)q->dtype, vv, &q->tab, q->tab.exact_cells, d_in, n_rec * rec_len, n_rec, rec_len, nullptr,
                                       d_tokens, out_stride, d_len, device, st);
        if (rc2 != ECGB_EUNSUPPORTED) return rc2;
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
    const VocabView *vv = ecgb_vocab_view(v);
    DeviceGuard g(device);
    uint8_t *d_text = nullptr; int32_t *d_tok = nullptr; int32_t *d_len = nullptr;
    const size_t stride = std::min(cap, n);
    cudaError_t e = cudaMalloc((void **)&d_text, n + 64);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_tok, std::max<size_t>(stride, 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_len, 4);
    int rc = ECGB_OK;
    unsigned long long count = 0;
    if (e == cudaSuccess) e = cudaMemcpy(d_text, h_text, n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        if (vv->compact && n >= 4096 && n < 0xFFFFFFF0ull && vv->max_token_len < 0xFFFFu) {
            // one long string: parallel over positions instead of one walker
            rc = ecgb_encode_long_device(v, d_text, n, reinterpret_cast<uint32_t *>(d_tok), stride, &count, 0);
        } else {
            int32_t len = 0;
            rc = ecgb_encode_symbols(v, d_text, 1, n, nullptr, d_tok, stride, d_len, nullptr);
            if (rc == ECGB_OK) e = cudaMemcpy(&len, d_len, 4, cudaMemcpyDeviceToHost);
            count = (unsigned long long)len;
        }
    }
    if (e == cudaSuccess && rc == ECGB_OK && h_out && stride)
        e = cudaMemcpy(h_out, d_tok, std::min<size_t>((size_t)count, stride) * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_text); cudaFree(d_tok); cudaFree(d_len);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? ECGB_ENOMEM : ECGB_ECUDA, "encode_text_host: %s", cudaGetErrorString(e));
    if (rc) return rc;
    *n_out = (size_t)count;
    if ((size_t)count > cap) return fail(ECGB_ECAPACITY, "output capacity %zu < %llu tokens", cap, count);
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
