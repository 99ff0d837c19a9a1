// E2: encode_text (reference: rust_bpe/src/lib.rs:149-193) fused with Q1
// (tokenizer_utils.py:14-19) for sm_100a -- the two-symbols-per-step walker.
//
// Semantics: greedy longest match over the trie -- at each token start walk as far as symbols
// match, remember the longest terminal seen, emit it and restart right after it
// (lib.rs:163-190).  This is NOT rank-ordered BPE merging.
//
// Mapping: one walker = one thread = one record, one CTA of up to 768 walkers per SM; the record
// axis alone fills the chip.  What a walker steps through is the PAIR TABLE (trie_host.h): a
// double-array automaton whose states are the trie nodes at even depth, so ONE 4-byte
// shared-memory gather advances a walker by TWO symbols (the round-1 kernel paid an 8-byte gather
// and a POPC rank per symbol; its l1tex pipe was 86 % busy).  A warp alternates between convergent
// phases:
//   refill : every lane with room streams the next 16 samples of ITS record with 256-bit loads
//            (LDG.E.ENL2.256: one full 32-byte sector per lane per request -- half the l1tex
//            wavefronts of 128-bit loads, profiles/micro/loadpat.cu), classifies them (K1's
//            arithmetic) and appends 16 PAIR CODES to its private ring in shared memory: entry p is
//            the table offset of the symbol pair (p-1, p), so a step needs one LDS.U16 and one add
//            to form the probe address.  Rings are transposed per warp (bank = lane), so ring
//            traffic is conflict-free wherever each cursor is.
//   walk   : bursts of kBurst branch-free pair steps (probe, compare the stored code, adopt the
//            next state; a lane whose probe fails just stops for the rest of the burst), then ONE
//            convergent block that serves every lane that failed: the odd-length probe
//            (first symbol alone), token id fetch, append to the lane's token queue, restart at
//            the root.  Tokens leave through the queue as 16-byte stores, eight per flush.
//   The end of a record is a sentinel class that matches nothing; record switches and bytes that are
//   their own token (text input) are handled at the phase boundary, outside the hot loop.  A ring
//   holds 64 symbols, of which at most ECGB_ENC_HIST (16) are history behind the cursor; a walk that
//   runs further than that past its last terminal (flat-line tokens of 64-256 symbols) lets the refill
//   overwrite its history, and if the token then ends before the oldest entry still in the ring, the
//   walker REWINDS: the ring restarts at the new token start and those symbols are read a second time.
//   Refill rounds that fewer than ECGB_ENC_MINLANES (20) lanes want are put off until they fill up,
//   unless a lane is running dry (a round costs the same for one lane as for 32).
// HBM traffic is the algorithmic minimum: samples once, tokens once.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "quant_device.cuh"

#ifndef ECGB_ENC_BURST
#define ECGB_ENC_BURST 4
#endif
#ifndef ECGB_ENC_RING
#define ECGB_ENC_RING 64      // ring capacity in symbols (a power of two)
#endif
#ifndef ECGB_ENC_HIST
#define ECGB_ENC_HIST 16  // symbols of history kept behind the cursor when the last terminal is further back (a longer
                          // look-back re-reads the record: rewind); unbounded history starves the lanes in long walks
#endif
#ifndef ECGB_ENC_MAXB
#define ECGB_ENC_MAXB 8
#endif
#ifndef ECGB_ENC_MINLANES
#define ECGB_ENC_MINLANES 20  // a refill round that fewer lanes want is put off unless one of them is short of symbols
#endif
#ifndef ECGB_ENC_NEEDY
#define ECGB_ENC_NEEDY 24     // "short of symbols": fewer than this many ahead of the cursor
#endif

namespace ecgb {

#ifdef ECGB_ENC_STATS
// instrumentation (profiles only): [0..8] walk phases by K, [16..48] refill rounds by lanes wanting,
// [50] rewinds, [51] failure blocks, [52] lanes served in failure blocks, [53] bursts, [54] lanes still running at burst end
__device__ unsigned long long g_enc_stats[64];
#define ENC_STAT(i, v) do { if (threadIdx.x == 0 && (blockIdx.x & 7u) == 0) atomicAdd(&g_enc_stats[i], (unsigned long long)(v)); } while (0)
#else
#define ENC_STAT(i, v) do { } while (0)
#endif

struct Enc2Args {
    const void *in;            // samples (or text bytes), all records back to back
    size_t n_total;            // total samples in `in`
    size_t n_rec, rec_len;
    const uint64_t *offsets;   // optional [n_rec + 1]
    int32_t *tokens;
    size_t out_stride;
    int32_t *lens;
    PairView pv;
    uint32_t vec_out;          // token rows are 16-byte aligned: flush with 128-bit stores
    uint32_t in_al32;          // `in` is 32-byte aligned: 256-bit loads
    QuantTables qt;
};

constexpr int kG2 = 16;          // symbols per refill
constexpr int kRing2 = ECGB_ENC_RING;  // ring capacity in symbols (2 bytes each)
constexpr int kHist = ECGB_ENC_HIST;
constexpr int kBurst = ECGB_ENC_BURST;
constexpr int kMaxBursts = ECGB_ENC_MAXB;  // bursts per walk phase (at most one token per burst: the queue of 16 cannot overflow)
static_assert(kMaxBursts <= 16 - 8, "token queue: 16 slots, flushed in eights");
constexpr uint32_t kRingWordMask = (uint32_t)(kRing2 / 2 - 1) << 7;  // word index bits of a ring address
constexpr int kThreads2 = 768;
constexpr uint32_t kRingWarpBytes = 32u * kRing2 * 2u;
constexpr uint32_t kQueueWarpBytes = 32u * 32u;
constexpr uint32_t kCheck = 0x3FFCu;  // the code field of a table entry

template <int DT> struct Elem2 { using T = typename SampleTraits<DT>::In; using Thr = typename SampleTraits<DT>::Thr; };
template <> struct Elem2<ECGB_U8> { using T = uint8_t; using Thr = float; };

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// ring entry of symbol position p, addressed by q2 = 2 * p: halfword (p & 63) of the lane's ring,
// 32-bit words transposed inside the warp's 4 KB-aligned region (word w of lane l at (w * 32 + l) * 4)
__device__ __forceinline__ uint32_t ring_at2(uint32_t ring_lane, int32_t q2) {
    return ring_lane + (((uint32_t)q2 & (uint32_t)(2 * kRing2 - 4)) << 5) + ((uint32_t)q2 & 2u);
}
// two entries further (one pair step): the word index lives in address bits 7..11
__device__ __forceinline__ uint32_t ring_next(uint32_t ra) {
    return ((ra + 128u) & kRingWordMask) | (ra & ~kRingWordMask);  // one LOP3
}
// token queue slot of token number t (16 halfwords per lane, transposed the same way)
__device__ __forceinline__ uint32_t queue_at(uint32_t queue_lane, uint32_t t) {
    return queue_lane + ((t & 14u) << 6) + ((t & 1u) << 1);
}

// 32 bytes at p (32-byte aligned when al32)
__device__ __forceinline__ void ldg32B(const void *p, bool al32, uint32_t (&r)[8]) {
    if (al32) {
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "l"(p));
    } else {
        const uint4 a = __ldg(static_cast<const uint4 *>(p)), b = __ldg(static_cast<const uint4 *>(p) + 1);
        r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
    }
}

// class << 2 of one sample (K1's arithmetic: quant_device.cuh)
template <int DT, bool CELLS>
__device__ __forceinline__ uint32_t class4_of(typename Elem2<DT>::T v, float lo, float scale, uint32_t cell_sa,
                                              const void *thr_smem, uint32_t cls_sa) {
    using Thr = typename Elem2<DT>::Thr;
    if constexpr (DT == ECGB_U8) {
        uint32_t c;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(c) : "r"(cls_sa + (uint32_t)v));
        return c;  // the staged table already holds class << 2
    } else if constexpr (CELLS) {
        float sf;
        const Thr s = to_thr(v, &sf);
        const uint32_t cell = cell_of(sf, lo, scale);
        Thr t;
        if constexpr (sizeof(Thr) == 4) {
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(cell_sa + cell * 4u));
        } else {
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"(cell_sa + cell * 8u));
        }
        return cell * 4u + (s >= t ? 4u : 0u);
    } else {
        float sf;
        const Thr s = to_thr(v, &sf);
        return classify_search<Thr>(s, static_cast<const Thr *>(thr_smem)) << 2;
    }
}

// 16 samples at global element index g (g % 16 == 0) -> 16 pair codes appended to the ring at symbol
// position p0 (p0 % 16 == 0).  Positions at or beyond `valid` are the sentinel.  prev = class << 2
// of the symbol before p0 (in/out).
template <int DT, bool CELLS>
__device__ __forceinline__ void refill16(const void *base, size_t g, size_t n_total, int valid, bool al32, uint32_t ring_lane,
                                         int32_t p0, uint32_t &prev, uint32_t wmul, uint32_t se4, float lo, float scale,
                                         uint32_t cell_sa, const void *thr_smem, uint32_t cls_sa) {
    using T = typename Elem2<DT>::T;
    constexpr int NB = sizeof(T) * 16 / 32;  // 32-byte pieces (u8: half a piece)
    T e[16];
    const T *p = static_cast<const T *>(base) + g;
    if (g + 16 <= n_total) {
        if constexpr (NB >= 1) {
            uint32_t r[NB][8];
#pragma unroll
            for (int j = 0; j < NB; j++) ldg32B(reinterpret_cast<const uint8_t *>(p) + 32 * j, al32, r[j]);
            memcpy(e, r, sizeof(e));
        } else {
            const uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
            memcpy(e, &r, sizeof(e));
        }
    } else {  // ragged end of the buffer
#pragma unroll
        for (int k = 0; k < 16; k++) e[k] = (g + k < n_total) ? p[k] : T(0);
    }
    uint32_t cs[16];
#pragma unroll
    for (int k = 0; k < 16; k++) cs[k] = class4_of<DT, CELLS>(e[k], lo, scale, cell_sa, thr_smem, cls_sa);
    if (valid < 16) {
#pragma unroll
        for (int k = 0; k < 16; k++)
            if (k >= valid) cs[k] = se4;
    }
    const uint32_t wa = ring_lane + (((uint32_t)p0 & (uint32_t)(kRing2 - 1)) << 6);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t d0 = prev * wmul + cs[2 * j];
        const uint32_t d1 = cs[2 * j] * wmul + cs[2 * j + 1];
        prev = cs[2 * j + 1];
        sts32(wa + 128u * j, d0 | (d1 << 16));
    }
}

template <int DT, bool CELLS, bool TOKS>
__global__ void __launch_bounds__(kThreads2, 1) encode2_kernel(Enc2Args a) {
    using Thr = typename Elem2<DT>::Thr;
    extern __shared__ __align__(16) uint8_t smem[];
    // layout: [pad to 4 KB] rings | queues | ent | tok | aux
    const uint32_t nwarps = blockDim.x >> 5;
    const uint32_t smem_sa = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t pad = (0u - smem_sa) & (kRingWarpBytes - 1u);
    uint8_t *s_rings = smem + pad;
    uint8_t *s_queues = s_rings + nwarps * kRingWarpBytes;
    const uint32_t n_ent = a.pv.n_ent;
    const uint32_t ent_bytes = (n_ent * 4u + 15u) & ~15u;
    const uint32_t tok_bytes = TOKS ? ((n_ent * 2u + 15u) & ~15u) : 0u;
    uint32_t *s_ent = reinterpret_cast<uint32_t *>(s_queues + nwarps * kQueueWarpBytes);
    uint16_t *s_tok = reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(s_ent) + ent_bytes);
    uint8_t *s_aux = reinterpret_cast<uint8_t *>(s_ent) + ent_bytes + tok_bytes;
    // aux: quantiser cell table + threshold list (sample dtypes) or the byte -> class << 2 table (text)
    QuantSmem<Thr> *qs = reinterpret_cast<QuantSmem<Thr> *>(s_aux);
    Thr *s_thr = reinterpret_cast<Thr *>(s_aux + sizeof(QuantSmem<Thr>));

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t tab_sa = (uint32_t)__cvta_generic_to_shared(s_ent);
    const uint32_t tok_sa = (uint32_t)__cvta_generic_to_shared(s_tok);
    const uint32_t cell_sa = (uint32_t)__cvta_generic_to_shared(qs);
    const uint32_t cls_sa = (uint32_t)__cvta_generic_to_shared(s_aux);
    const uint32_t ring_lane = smem_sa + pad + warp * kRingWarpBytes + lane * 4u;
    const uint32_t queue_lane = (uint32_t)__cvta_generic_to_shared(s_queues) + warp * kQueueWarpBytes + lane * 4u;

    for (uint32_t i = threadIdx.x; i < n_ent; i += blockDim.x) s_ent[i] = a.pv.d_ent[i];
    if (TOKS)
        for (uint32_t i = threadIdx.x; i < n_ent; i += blockDim.x) s_tok[i] = a.pv.d_tok[i];
    if constexpr (DT == ECGB_U8) {
        for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_aux[i] = (uint8_t)(a.pv.d_cls[i] << 2);
    } else {
        load_quant_smem(qs, a.qt);
        if (threadIdx.x < kNumThresholds) s_thr[threadIdx.x] = static_cast<const Thr *>(a.qt.d_thr)[threadIdx.x];
    }
    // rings start out as zeros: a stale entry is always a harmless table offset
    for (uint32_t i = threadIdx.x; i < nwarps * (kRingWarpBytes + kQueueWarpBytes) / 4u; i += blockDim.x)
        reinterpret_cast<uint32_t *>(s_rings)[i] = 0u;
    __syncthreads();

    const float qlo = a.qt.lo, qscale = a.qt.scale;
    const uint32_t W = a.pv.W;
    const uint32_t wmul = 1u << W;
    const uint32_t sm_code = a.pv.SM << 2;  // low field of the odd-length ("single") probe
    const uint32_t se4 = a.pv.SE << 2;      // sentinel class, scaled
    const uint32_t himask = (wmul - 1u) << (W + 2u);
    const uint32_t lomask = (wmul - 1u) << 2;
    const uint32_t rootA = tab_sa + a.pv.root_base * 4u;
    const uint32_t stride = a.out_stride < 0x7fffffffu ? (uint32_t)a.out_stride : 0x7fffffffu;
    const bool al32 = a.in_al32 != 0;
    const bool vec_out = a.vec_out != 0;
    const size_t n_total = a.offsets ? (size_t)a.offsets[a.n_rec] : a.n_total;
    constexpr unsigned FULL = 0xffffffffu;

    // contiguous, even split of the records over the CTAs
    const size_t r_lo = (size_t)(((unsigned __int128)a.n_rec * blockIdx.x) / gridDim.x);
    const size_t r_hi = (size_t)(((unsigned __int128)a.n_rec * (blockIdx.x + 1)) / gridDim.x);
    size_t r_next = r_lo + threadIdx.x;

    // walker state; positions are relative to `org` (a multiple of the refill group)
    bool active = false, done = false, parked = false, rewind = false, closed = false;
    size_t org = 0, r_cur = 0;
    int32_t end32 = 0;   // record end
    int32_t hi32 = 0;    // the ring holds the entries of positions [hi32 - kRing2, hi32)
    int32_t q = 0;       // 2 * position of the next symbol to read
    int32_t mq = 0;      // 2 * position the ring should keep: start of the pair that set `m`, or of the walk
    uint32_t A = rootA;  // shared-memory address of the current state's row
    uint32_t ra = ring_lane;  // ring address of the entry at the cursor (pair (q/2, q/2 + 1))
    uint32_t cw = 0;     // ... and the entry
    uint32_t m = 0;      // last probe that passed a terminal: probe address << 8 | low byte of its entry
    uint32_t cnt = 0, flushed = 0, prev = se4;
    int32_t *outp = nullptr;

    auto tok_of = [&](uint32_t ref) -> uint32_t {  // token id of the slot at shared address `ref`
        const uint32_t off2 = (ref - tab_sa) >> 1;
        if (TOKS) return lds16(tok_sa + off2);
        return (uint32_t)__ldg(a.pv.d_tok + (off2 >> 1));
    };
    auto drain = [&]() {  // every queued token, one by one (record ends, rare paths)
        for (uint32_t t = flushed; t < cnt; t++)
            if (t < stride) outp[t] = (int32_t)lds16(queue_at(queue_lane, t));
        flushed = cnt;
    };

    for (;;) {
        // ------------------------------------------------ phase boundary (convergent)
        // (1) parked walkers
        if (parked) {
            const int32_t pos = q >> 1;
            if (rewind) {
                // the token ended before the oldest entry still in the ring: restart the ring there
                const int32_t shift = pos & ~(kG2 - 1);
                org += (size_t)shift;
                end32 -= shift;
                q = mq = 2 * (pos - shift);
                hi32 = 0;
                closed = false;
                prev = se4;
                rewind = false;
            } else {
                // end of the record, or a byte that occurs in no merge and is its own token
                // (text only; lib.rs:155-157)
                drain();
                if (pos >= end32) {
                    a.lens[r_cur] = (int32_t)cnt;
                    active = false;
                } else {
                    const uint32_t byte = static_cast<const uint8_t *>(a.in)[org + (size_t)pos];
                    if (cnt < stride) outp[cnt] = (int32_t)byte;
                    cnt++;
                    flushed = cnt;
                    q += 2;
                    mq = q;
                    m = 0;
                    A = rootA;
                }
            }
            parked = false;
        }
        // (2) next record
        if (!active && !done) {
            if (r_next < r_hi) {
                r_cur = r_next;
                r_next += blockDim.x;
                const size_t rs = a.offsets ? (size_t)a.offsets[r_cur] : r_cur * a.rec_len;
                const size_t re = a.offsets ? (size_t)a.offsets[r_cur + 1] : rs + a.rec_len;
                org = rs & ~(size_t)(kG2 - 1);
                end32 = (int32_t)(re - org);
                q = mq = 2 * (int32_t)(rs - org);
                hi32 = 0;
                closed = false;
                A = rootA;
                m = 0;
                cnt = flushed = 0;
                prev = se4;
                outp = a.tokens + r_cur * a.out_stride;
                active = true;
            } else {
                done = true;
                A = rootA;
                cw = 0;
            }
        }
        if (__all_sync(FULL, done)) break;
        // (3) token queues: eight tokens at a time as two 16-byte stores
        {
            const bool fl = active && cnt - flushed >= 8u;
            if (__any_sync(FULL, fl)) {
                if (fl) {
                    const uint32_t qa = queue_lane + ((flushed & 8u) << 6);
                    const uint32_t w0 = lds32(qa), w1 = lds32(qa + 128u), w2 = lds32(qa + 256u), w3 = lds32(qa + 384u);
                    if (vec_out && flushed + 8u <= stride) {
                        int4 *dst = reinterpret_cast<int4 *>(outp + flushed);
                        dst[0] = make_int4((int)(w0 & 0xFFFFu), (int)(w0 >> 16), (int)(w1 & 0xFFFFu), (int)(w1 >> 16));
                        dst[1] = make_int4((int)(w2 & 0xFFFFu), (int)(w2 >> 16), (int)(w3 & 0xFFFFu), (int)(w3 >> 16));
                    } else {
                        const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
                        for (int k = 0; k < 8; k++)
                            if (flushed + k < stride) outp[flushed + k] = (int32_t)((w[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu);
                    }
                    flushed += 8u;
                }
            }
        }
        // (4) refill, 16 symbols at a time: while the ring still keeps everything from mq on, or --
        //     when the walker cannot run a burst otherwise -- over its own history (see `rewind`)
#pragma unroll 1
        for (int g = 0; g < kRing2 / kG2; g++) {
            const bool room = hi32 - max(mq >> 1, (q >> 1) - kHist) <= kRing2 - kG2 - 2;
            const bool starved = hi32 - (q >> 1) < 2 * kBurst;
            const bool want = active && !closed && (room || starved);
            // A round costs the same whether one lane or all of them take part, and every lane needs 1 group per 16
            // symbols it consumes whatever the schedule: rounds that only a few lanes want (a lane whose consumption has
            // just crossed a multiple of 16) are put off until they fill up -- unless such a lane is running dry, which
            // would cut the next walk phase short for the whole warp.
            {
                const unsigned wm = __ballot_sync(FULL, want);
                if (!wm) break;
                const bool needy = want && (hi32 - (q >> 1) < ECGB_ENC_NEEDY);
#ifdef ECGB_ENC_NEEDY_FIRST_ROUND_ONLY
                if (__popc(wm) < ECGB_ENC_MINLANES && (g > 0 || !__any_sync(FULL, needy))) break;
#else
                if (__popc(wm) < ECGB_ENC_MINLANES && !__any_sync(FULL, needy)) break;
#endif
            }
#ifdef ECGB_ENC_STATS
            { const int nw_ = __popc(__ballot_sync(FULL, want)); ENC_STAT(16 + nw_, 1); }
#endif
            if (want) {
                refill16<DT, CELLS>(a.in, org + (size_t)hi32, n_total, end32 - hi32, al32, ring_lane, hi32, prev, wmul, se4, qlo,
                                    qscale, cell_sa, s_thr, cls_sa);
                hi32 += kG2;
                if (hi32 >= end32) {
                    // the record is complete: two entries of sentinel close it (a walker never steps
                    // past them, so nothing more is needed)
                    sts32(ring_lane + (((uint32_t)hi32 & (kRing2 - 1)) << 6), (prev * wmul + se4) | ((se4 * wmul + se4) << 16));
                    hi32 += 2;
                    closed = true;
                }
            }
        }
        // ------------------------------------------------ walk
        uint32_t nb = 0x7fffffffu;
        if (active && !closed) nb = (uint32_t)(hi32 - (q >> 1)) / (2u * kBurst);
        uint32_t K = min(__reduce_min_sync(FULL, nb), (uint32_t)kMaxBursts);
        ENC_STAT(K, 1);
        ra = ring_at2(ring_lane, q + 2);
        cw = lds16(ra);
        bool walk = active;
        for (; K > 0; K--) {
            bool run = walk;
#pragma unroll
            for (int j = 0; j < kBurst; j++) {
                const uint32_t addr = A + cw;
                const uint32_t e = lds32(addr);
                const uint32_t ran = ring_next(ra);
                const bool ok = run && ((e ^ cw) & kCheck) == 0u;
                if (ok && (e & 3u) != 0u) { m = __byte_perm(e, addr, 0x6540); mq = q; }
                if (ok) { A = tab_sa + (e >> 14); q += 4; ra = ran; cw = lds16(ran); }
                run = ok;
            }
            const bool f = walk && !run;
#ifdef ECGB_ENC_STATS
            { const int nr_ = __popc(__ballot_sync(FULL, run)), nf_ = __popc(__ballot_sync(FULL, f));
              ENC_STAT(53, 1); ENC_STAT(54, nr_); if (nf_) { ENC_STAT(51, 1); ENC_STAT(52, nf_); } }
#endif
            if (__any_sync(FULL, f)) {
                if (f) {
                    // the pair at the cursor is no edge: the first symbol alone may still reach a token
                    const uint32_t x1 = (cw & himask) | sm_code;
                    const uint32_t addr1 = A + x1;
                    const uint32_t e1 = lds32(addr1);
                    // which slot holds the token and where the next one starts, without branches (the four cases would
                    // otherwise run one after the other inside this already thinly populated block): the single probe
                    // hit, else the pair that set m ended on a terminal, else only its first symbol did
                    const bool hit1 = ((e1 ^ x1) & kCheck) == 0u;
                    const bool m2 = (m & 2u) != 0u;
                    const uint32_t ref_m = (m >> 8) - (m2 ? 0u : (m & lomask) - sm_code);
                    const uint32_t ref = hit1 ? addr1 : ref_m;
                    const int32_t nq = hit1 ? q + 2 : mq + (m2 ? 4 : 2);
                    const bool got = hit1 || (m & 3u) != 0u;  // else: sentinel at the root -- end of record, or a byte that is its own token
                    // the rest without branches too: lanes that got a token queue it and restart at the root; a lane whose
                    // restart point has left the ring (left) or that got nothing parks until the phase boundary
                    const bool left = got && ((nq >> 1) + 1 < hi32 - kRing2);
                    const uint32_t tok = tok_of(got ? ref : tab_sa);
                    if (got) sts16(queue_at(queue_lane, cnt), tok);
                    cnt += got ? 1u : 0u;
                    q = got ? nq : q;
                    mq = got ? nq : mq;
                    m = got ? 0u : m;
                    A = got ? rootA : A;
                    parked = parked || !got || left;
                    rewind = rewind || left;
                    walk = got && !left;
#ifdef ECGB_ENC_STATS
                    if (left && (blockIdx.x & 7u) == 0 && threadIdx.x < 32) atomicAdd(&g_enc_stats[50], 1ull);
#endif
                    ra = ring_at2(ring_lane, q + 2);  // harmless for a parked lane: any ring address is a valid one
                    cw = lds16(ra);
                }
            }
        }
    }
}

template <int DT>
static int launch_encode2_t(const Enc2Args &a, int exact_cells, int device, cudaStream_t st) {
    using Thr = typename Elem2<DT>::Thr;
    const int sms = sm_count(device);
    int smem_max = 0;
    ECGB_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    // tables + quantiser cells (or the byte -> class table) + slack to align the rings to 4 KB
    const size_t aux = (DT == ECGB_U8 ? 256 : ((sizeof(QuantSmem<Thr>) + sizeof(Thr) * 32 + 15) & ~(size_t)15)) + kRingWarpBytes;
    const size_t ent_bytes = ((size_t)a.pv.n_ent * 4 + 15) & ~(size_t)15;
    const size_t tok_bytes = ((size_t)a.pv.n_ent * 2 + 15) & ~(size_t)15;
    const size_t per_warp = kRingWarpBytes + kQueueWarpBytes;
    if ((size_t)smem_max < ent_bytes + aux + per_warp) return ECGB_EUNSUPPORTED;  // caller falls back (no message)
    // one walker per record; CTAs get equal contiguous record ranges
    size_t grid = std::min<size_t>((size_t)sms, (a.n_rec + 31) / 32);
    if (grid < 1) grid = 1;
    const size_t per_cta = (a.n_rec + grid - 1) / grid;
    static const int knob = getenv("ECGB_ENC_WARPS") ? atoi(getenv("ECGB_ENC_WARPS")) : 0;  // tuning knob
    const size_t warps_max = knob > 0 ? (size_t)knob : (size_t)kThreads2 / 32;
    // token ids in shared memory when that leaves at least 16 warps of walkers, else in L2
    const size_t fit_tok = (size_t)smem_max > ent_bytes + tok_bytes + aux ? ((size_t)smem_max - ent_bytes - tok_bytes - aux) / per_warp : 0;
    const size_t fit_notok = ((size_t)smem_max - ent_bytes - aux) / per_warp;
    // ... unless leaving them in L2 saves a whole pass over the CTA's records (more walkers at once)
    static const char *tok_knob = getenv("ECGB_ENC_TOKS");  // A/B: 0 = ids in L2, 1 = ids in shared memory when they fit
    const size_t want_warps = std::min(warps_max, (per_cta + 31) / 32);
    auto passes_with = [&](size_t fit) { const size_t c = std::min(warps_max, fit); return c ? (per_cta + c * 32 - 1) / (c * 32) : (size_t)1 << 30; };
    bool toks = fit_tok >= std::min<size_t>(16, warps_max);
    if (toks && fit_tok < want_warps && passes_with(fit_notok) < passes_with(fit_tok)) toks = false;
    if (tok_knob) toks = atoi(tok_knob) != 0 && fit_tok >= 1;
    const size_t cap = std::min(warps_max, toks ? fit_tok : fit_notok);
    if (cap < 1) return ECGB_EUNSUPPORTED;
    // as few passes over the CTA's records as the walker limit allows, lanes spread evenly over them
    const size_t passes = (per_cta + cap * 32 - 1) / (cap * 32);
    const size_t w = std::max<size_t>(1, std::min(cap, ((per_cta + passes - 1) / passes + 31) / 32));
    const size_t smem = ent_bytes + (toks ? tok_bytes : 0) + aux + w * per_warp;
    auto kern = exact_cells ? (toks ? encode2_kernel<DT, true, true> : encode2_kernel<DT, true, false>)
                            : (toks ? encode2_kernel<DT, false, true> : encode2_kernel<DT, false, false>);
    ECGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, (unsigned)(w * 32), smem, st>>>(a);
    ECGB_CUDA(cudaGetLastError());
#ifdef ECGB_ENC_STATS
    {
        unsigned long long h[64], z[64] = {0};
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_enc_stats, sizeof(h));
        cudaMemcpyToSymbol(g_enc_stats, z, sizeof(z));
        fprintf(stderr, "enc2 stats: K:");
        for (int i = 0; i <= 8; i++) fprintf(stderr, " %llu", h[i]);
        fprintf(stderr, "\n  refill rounds by lanes wanting (1..32):");
        for (int i = 1; i <= 32; i++) fprintf(stderr, " %llu", h[16 + i]);
        fprintf(stderr, "\n  rewinds %llu, failure blocks %llu serving %llu lanes, bursts %llu with %llu lanes still running at the end\n",
                h[50], h[51], h[52], h[53], h[54]);
    }
#endif
    return ECGB_OK;
}

// Fused quantise + encode (dt = sample type) or encode of text bytes (dt = ECGB_U8) through the pair
// table.  ECGB_EUNSUPPORTED (without a message) when the vocabulary has no pair table or it does
// not fit in shared memory: the caller then takes the bitmap-trie kernel.
int launch_encode2(int dt, const VocabView *vv, const QuantTables *qt, int exact_cells, const void *d_in, size_t n_total,
                   size_t n_rec, size_t rec_len, const uint64_t *d_offsets, int32_t *d_tokens, size_t out_stride,
                   int32_t *d_len, int device, cudaStream_t st) {
    if (!vv->pair.d_ent) return ECGB_EUNSUPPORTED;
    if (rec_len >= (1ull << 30)) return ECGB_EUNSUPPORTED;
    // Which walker (measured on B200, profiles/encode_ab.py, profiles/encode_cases.py):
    //   * the 8-byte trie does not fit in shared memory next to 768 walkers' rings (10 000 merges): this kernel, whose
    //     pair table is 2.6x smaller -- 14.3 ms against 27.4 ms for 100 k records;
    //   * the trie fits and the batch fills the chip (>= 512 records per SM): this kernel for float samples -- 13.31
    //     against 13.34 ms (fp32), 6.67 against 6.72 (12 x 2500), both kernels with deferred refill rounds; the bitmap
    //     kernel for int16 samples (12.7 against 13.0 ms);
    //   * smaller batches: the bitmap kernel, whose step chain is shorter -- a launch is then as long as ONE record's walk
    //     (2 048 records: 8.0 against 9.5 ms; 2 records of 12 x 500: 0.54 against 0.63 ms).
    // A/B knobs: ECGB_ENCODE_V1 / ECGB_ENCODE_V2 force one.
    static const bool force_v1 = getenv("ECGB_ENCODE_V1") != nullptr, force_v2 = getenv("ECGB_ENCODE_V2") != nullptr;
    if (force_v1) return ECGB_EUNSUPPORTED;
    if (!force_v2) {
        int smem_max = 0;
        ECGB_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        const size_t trie_and_rings = (size_t)vv->n_nodes * 8 + 768 * 128 + 4096;
        const bool trie_fits = trie_and_rings <= (size_t)smem_max;
        if (trie_fits && n_rec < (size_t)512 * (size_t)sm_count(device)) return ECGB_EUNSUPPORTED;
        if (trie_fits && dt == ECGB_I16) return ECGB_EUNSUPPORTED;  // int16 full batch: 12.7 ms (bitmap) against 13.0 ms
    }
    Enc2Args a{};
    a.in = d_in; a.n_total = n_total; a.n_rec = n_rec; a.rec_len = rec_len; a.offsets = d_offsets;
    a.tokens = d_tokens; a.out_stride = out_stride; a.lens = d_len;
    a.pv = vv->pair;
    a.vec_out = (((uintptr_t)d_tokens & 15) == 0 && (out_stride & 3) == 0) ? 1u : 0u;
    a.in_al32 = ((uintptr_t)d_in & 31) == 0 ? 1u : 0u;
    if (qt) a.qt = *qt;
    switch (dt) {
        case ECGB_F32: return launch_encode2_t<ECGB_F32>(a, exact_cells, device, st);
        case ECGB_F64: return launch_encode2_t<ECGB_F64>(a, exact_cells, device, st);
        case ECGB_I16: return launch_encode2_t<ECGB_I16>(a, exact_cells, device, st);
        case ECGB_U8: return launch_encode2_t<ECGB_U8>(a, 1, device, st);
    }
    return ECGB_EUNSUPPORTED;
}

}  // namespace ecgb
