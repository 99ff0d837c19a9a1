// T1-T4: byte_pair_encoding (reference: rust_bpe/src/lib.rs:58-125) on sm_100a.
//
// Reference loop per merge step: get_stats (full recount of overlapping pairs into a
// hash map, lib.rs:28-48) -> argmax (lib.rs:92-94) -> merge (single-threaded in-place
// left-to-right replacement, lib.rs:10-26).
//
// Here, per step, ONE streaming pass over the token stream (2-byte ids, read once, written once -- the algorithmic
// 2*(n_t + n_{t+1}) bytes of SURVEY.md 8d):
//   * merge_pass: tiles of 4096 tokens; every merge site also patches the pair histogram for the windows it
//     destroys/creates (block-private shared-memory table, flushed once per CTA into an open-addressing hash table in
//     global memory/L2), so the histogram always equals a full recount and no second pass over the tokens is needed;
//   * argmax with the deterministic tie rule (max count, then smallest (left,right)) over a candidate list (every slot
//     whose count is >= tau), falling back to a table scan that lowers tau when the list runs dry.
// One device: train_loop_kernel runs every step inside ONE persistent cooperative kernel.  The stream is cut into one
// contiguous CHUNK per CTA before the first step (ChunkWhere / chunk_step_begin); a chunk is merged in place -- in the
// first token buffer while it is longer than 12 288 tokens, in the CTA's shared memory afterwards -- and a CTA treats the
// other CTAs exactly like the ranks of a sharded run: all it needs from them is their 64-byte boundary record.  (The
// round-1 scheme -- one stream, ping-pong buffers, decoupled look-back over tile tickets -- remains for the step-wise
// merge_kernel and behind ECGB_RESIDENT_TAIL=1.)
// (x,x) pairs: merge() is greedy left to right, so inside a run of x only the elements at even run offsets start a
// site (lib.rs:14-18); the run position is carried from tile to tile and from chunk to chunk (run parity in the records).
//
// Sharded corpora (one trainer per rank = GPU): dist_loop_kernel is the same persistent loop on every rank;
// the chain of chunks continues into the neighbouring ranks through one shard record per peer, every rank keeps a copy
// of the GLOBAL histogram and takes the same argmax, and the ranks exchange histogram patches and shard records by
// writing self-validating 8-byte units straight into each other's memory over NVLink (PeerView; no collective, no host,
// no fence inside the loop).  The step-wise ecgb_trainer_dist_* calls (launches + two host-enqueued all-gathers per
// step) remain as the NCCL reference path and provide the initial records / get_stats exchange.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.h"

namespace cg = cooperative_groups;

namespace ecgb {

constexpr uint32_t kEmptyKey = 0xFFFFFFFFu;
constexpr uint32_t kSentinel = 0xFFFFu;  // never a token id (ids <= 0xFFFE)
constexpr int kTPB = 256;                // threads per block of the merge kernel
constexpr int kIPT = 16;                 // tokens per thread
constexpr int kTile = kTPB * kIPT;
constexpr int kBoundaryWords = 16;       // u32 words of boundary info per rank
constexpr int kCtasPerSm = 4;            // co-resident CTAs per SM of the persistent kernels (launch bounds)
constexpr int kMaxWorld = 16;            // ranks of one sharded run (the GPUs of one box)
constexpr int kApplyCtas = 16;           // persistent sharded loop: CTAs that apply one peer's patch list
constexpr uint32_t kRedundantArgmax = 2048;  // candidate lists up to this size are scanned by every CTA
constexpr int kChunkTiles = 3;           // resident tail: each CTA keeps up to 3 tiles of the stream in shared memory
constexpr int kChunkCap = kChunkTiles * kTile;

struct PairTable {
    uint32_t *keys;            // left << 16 | right, kEmptyKey = free
    unsigned long long *cnt;   // two's-complement adds; always >= 0 between steps
    uint32_t mask;             // capacity - 1
    uint32_t *used;            // slots claimed
    uint32_t *overflow;        // set when a probe sequence wrapped
    // argmax candidates: every slot whose count is >= *tau is listed in cand[0 .. *ncand)
    // (inbits marks listed slots).  NULL for tables that are never arg-maxed.
    const unsigned long long *tau;
    uint32_t *cand, *ncand, *inbits;
    // persistent sharded loop: every add to this view of the table is also appended to the rank's
    // outgoing patch list in the peers' memory (NULL otherwise)
    struct PeerPush *push;
};

// Outgoing patch list of one merge step (persistent sharded loop): entry i goes to slot i of this
// rank's inbox on every peer, written there directly over NVLink as one self-validating 16-byte unit.
struct PeerPush {
    uint32_t *out_count;       // (local) entries appended so far in this step
    uint32_t *overflow;        // (local) set when the inbox capacity is exceeded
    uint32_t cap;
    uint32_t tag;              // tag of the step (peer_tag)
    int rank, world;
    uint4 *dst[kMaxWorld];     // dst[r]: this rank's inbox (for this step's parity) in rank r's memory
};

struct Best {
    unsigned long long count;  // 0 = no pair
    uint32_t key;
    uint32_t ntied;
};

struct Boundary {  // one per rank, kBoundaryWords u32
    uint32_t n_lo, n_hi;     // shard length
    uint32_t first[3];       // first min(3, n) tokens
    uint32_t last[2];        // last[1] = final token, last[0] = the one before (if n >= 2)
    uint32_t trail_par;      // parity of the trailing run of best.left (x,x steps)
    uint32_t all_a;          // the whole shard is a run of best.left
    uint32_t pad[7];
};
static_assert(sizeof(Boundary) == kBoundaryWords * 4, "boundary layout");

struct Halo {
    uint32_t nl, nr;         // usable left / right context tokens
    uint32_t L[2];           // L[1] is adjacent to the shard's first token
    uint32_t R[3];           // R[0] is adjacent to the shard's last token
    uint32_t par_in;         // parity of the run of `left` ending just before the shard
};

struct DevState {
    unsigned long long n[2];  // token count of buffer 0 / 1
    uint32_t done_step;       // first step that found no pair (0xFFFFFFFF = none)
    uint32_t argmax_done;     // block completion counter of argmax_kernel (self-resetting)
    uint32_t cur_step;        // device-side step counter (ECGB_STEP_DEVICE: whole steps replayed from a CUDA graph)
    uint32_t max_steps;       // = max_merges: steps beyond it are ignored
};
constexpr uint32_t kStepFromDevice = 0xFFFFFFFFu;
static_assert(kStepFromDevice == ECGB_STEP_DEVICE, "header constant");

struct TrainView {
    uint16_t *tok[2];
    DevState *dev;
    PairTable main, delta;
    Best *best;               // [max_merges + 1]
    Best *partial;            // [cta_stride + 1]
    uint32_t *tickets;        // [max_merges + 1] tile tickets, zero-initialised
    unsigned long long *tile_status;  // [max tiles] decoupled look-back (merge_kernel)
    Boundary *boundary;       // this rank's boundary info (device)
    unsigned long long *n_hist;  // [max_merges + 2] stream length before each step
    unsigned int *arrive;     // train_loop_kernel: CTAs that have finished the argmax of a step, cumulative
    Boundary *cta_bd;         // [2][cta_stride] resident tail: boundary record of every CTA's chunk, by step parity
    uint32_t *cta_counts;     // [cta_stride] resident tail: chunk lengths for the final write-back
    uint32_t resident_ok;     // resident tail enabled
    uint32_t redundant_max;   // train_loop_kernel: candidate lists up to this size are scanned by every CTA
    uint32_t cta_stride;      // CTA slots of partial / cta_bd / cta_counts (= SMs x kCtasPerSm)
    const uint32_t *new_ids;  // id created by step i (NULL: 256 + i, lib.rs:97); set by ecgb_trainer_apply_pairs
    unsigned int *gbar;       // persistent sharded loop: arrival counter of its grid barrier (cumulative)
    unsigned int *abort;      // persistent sharded loop: set when a wait timed out; every spin loop gives up
    int rank, world;
};

// ---- device-initiated exchange of the persistent sharded loop (dist_loop_kernel) ----
// Every rank owns a receive AREA in its own memory; peers map it (CUDA IPC between the one-process-per-GPU
// ranks, or plain pointers inside one process) and write into it directly over NVLink.  Everything a peer
// writes is SELF-VALIDATING (the idea of NCCL's LL protocol): each 8-byte unit carries, next to 4 bytes of
// payload, the tag of the step it belongs to, and 8-byte stores arrive whole -- so neither side needs a
// fence (a system-scope fence costs ~5 us here; three per step were most of the step time of the first
// version of this exchange).  tag = epoch << 20 | (step + 1); the epoch changes with every run, so units
// left over from an earlier run never validate.
//   flag[parity][src]     64 B  unit 0: (entries of src's patch list of the step, tag)
//   rec [parity][src]     64 B  units 0-4: src's shard record of the stream the step with that tag READS --
//                               (n_lo), (n_hi), (first0 | first1 << 16), (first2 | last0 << 16), (last1);
//                               unit 5: its (x,x) run fields (trail_par | all_a << 1), second exchange of the step
//   ent [parity][src][cap] 16 B (key, tag, delta, tag)
// Two parities: a rank can be at most one step ahead of the slowest peer (it needs every peer's flag of step
// t before its argmax of step t + 1).
struct PeerView {
    int rank, world;
    uint32_t cap;                  // list entries per (parity, source)
    uint32_t epoch;                // run number (see tag)
    uint8_t *area[kMaxWorld];      // area[r] = rank r's area as mapped here; area[rank] is this rank's own
    uint32_t *out_count;           // (local) [2] entries of this rank's outgoing list, by step parity
    uint32_t *done_ctas;           // (local) CTAs that have flushed their patches, cumulative over steps
    uint32_t *overflow;            // (local) outgoing list overflow
    unsigned long long timeout_ns; // a wait for a peer gives up after this long (sets *abort)
};
__host__ __device__ inline size_t peer_flag_off(int world, int parity, int src) { return ((size_t)parity * world + src) * 64; }
__host__ __device__ inline size_t peer_rec_off(int world, int parity, int src) { return ((size_t)(2 + parity) * world + src) * 64; }
__host__ __device__ inline size_t peer_ent_off(int world, uint32_t cap, int parity, int src) {
    return (size_t)4 * world * 64 + ((size_t)parity * world + src) * (size_t)cap * 16;
}
__host__ __device__ inline size_t peer_area_bytes(int world, uint32_t cap) { return peer_ent_off(world, cap, 2, 0); }
__host__ __device__ inline uint32_t peer_tag(uint32_t epoch, uint32_t step) { return (epoch << 20) | ((step + 1u) & 0xFFFFFu); }
__host__ __device__ inline unsigned long long peer_unit(uint32_t payload, uint32_t tag) { return (unsigned long long)payload | ((unsigned long long)tag << 32); }
// payloads of record units 2-4

__device__ __forceinline__ uint32_t hash_key(uint32_t k) {
    k ^= k >> 16;
    k *= 0x7feb352dU;
    k ^= k >> 15;
    k *= 0x846ca68bU;
    k ^= k >> 16;
    return k;
}

// tau_val: the candidate threshold (*t.tau), read by the caller once -- it only changes between passes,
// and loading it here would put one more L2 round trip behind every atomic
__device__ __forceinline__ void push_entry(const PeerPush &p, uint32_t key, int delta) {
    // one atomic per converged group of callers
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(p.out_count, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    const uint32_t idx = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
    if (idx >= p.cap) { *p.overflow = 1u; return; }
    const uint4 e = make_uint4(key, p.tag, (uint32_t)delta, p.tag);
    for (int r = 0; r < p.world; r++)
        if (r != p.rank) p.dst[r][idx] = e;
}

__device__ __forceinline__ void table_add(const PairTable &t, uint32_t key, long long delta, unsigned long long tau_val) {
    if (t.push != nullptr) push_entry(*t.push, key, (int)delta);
    uint32_t slot = hash_key(key) & t.mask;
    // a probe sequence this long means the table is (nearly) full: give up and report ECGB_ECAPACITY instead of crawling
    const uint32_t max_probes = min(t.mask, 4095u);
    for (uint32_t probes = 0; probes <= max_probes; probes++) {
        uint32_t k = t.keys[slot];
        if (k == kEmptyKey) {
            uint32_t old = atomicCAS(&t.keys[slot], kEmptyKey, key);
            if (old == kEmptyKey) { atomicAdd(t.used, 1u); k = key; } else { k = old; }
        }
        if (k == key) {
            const unsigned long long old = atomicAdd(&t.cnt[slot], (unsigned long long)delta);
            if (delta > 0 && t.tau != nullptr && old + (unsigned long long)delta >= tau_val) {
                const uint32_t bit = 1u << (slot & 31);
                if (!(atomicOr(&t.inbits[slot >> 5], bit) & bit)) t.cand[atomicAdd(t.ncand, 1u)] = slot;
            }
            return;
        }
        slot = (slot + 1) & t.mask;
    }
    atomicExch(t.overflow, 1u);
}

// Block-private patch table: the histogram patches of one merge pass are first folded in
// shared memory (early merge steps hit a few hundred keys millions of times) and flushed
// to the global table once per CTA per step.
constexpr int kPatchBits = 10;
constexpr int kPatchSlots = 1 << kPatchBits;
struct PatchTable {
    uint32_t keys[kPatchSlots];
    int vals[kPatchSlots];
    // persistent kernels: the global table may only change once every CTA has taken this step's argmax
    // from it (gate counter >= gate_target); NULL when launches order the two
    unsigned int *gate;
    unsigned int gate_target;
};

__device__ __forceinline__ void table_add(const PairTable &t, uint32_t key, long long delta) {
    table_add(t, key, delta, t.tau != nullptr ? *t.tau : ~0ull);
}

__device__ __forceinline__ void patch_add(PatchTable &p, const PairTable &t, uint32_t key, int delta) {
    uint32_t slot = (key * 0x9E3779B1u) >> (32 - kPatchBits);  // multiplicative hash: the table is private
#pragma unroll 1
    for (int probes = 0; probes < 16; probes++) {
        uint32_t k = p.keys[slot];
        if (k == kEmptyKey) {
            const uint32_t old = atomicCAS(&p.keys[slot], kEmptyKey, key);
            k = old == kEmptyKey ? key : old;
        }
        if (k == key) {
            atomicAdd(&p.vals[slot], delta);
            return;
        }
        slot = (slot + 1) & (kPatchSlots - 1);
    }
    // private table crowded: go to the global one
    if (p.gate != nullptr) {
        unsigned int spins = 0;
        while (*reinterpret_cast<volatile unsigned int *>(p.gate) < p.gate_target)
            if (++spins > (1u << 28)) break;
    }
    table_add(t, key, (long long)delta);
}

__device__ __forceinline__ void patch_clear(PatchTable &p) {
    for (int i = threadIdx.x; i < kPatchSlots; i += blockDim.x) { p.keys[i] = kEmptyKey; p.vals[i] = 0; }
}

__device__ __forceinline__ void patch_flush(PatchTable &p, const PairTable &t, unsigned long long tau_val) {
    for (int i = threadIdx.x; i < kPatchSlots; i += blockDim.x)
        if (p.keys[i] != kEmptyKey && p.vals[i] != 0) table_add(t, p.keys[i], (long long)p.vals[i], tau_val);
}

__device__ __forceinline__ uint32_t mk(uint32_t l, uint32_t r) { return (l << 16) | r; }

// Optional phase timing of one CTA of train_loop_kernel (build with ECGB_NVCC_EXTRA=-DECGB_TRAIN_TIMING,
// run profiles/train_phases.py): thread 0 of CTA 1 accumulates the time between marks.
#ifdef ECGB_TRAIN_TIMING
__device__ unsigned long long g_phase[16];
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define ECGB_MARK(i)                                                       \
    do {                                                                   \
        if (blockIdx.x == 1 && threadIdx.x == 0) {                         \
            const unsigned long long t_ = now_ns();                        \
            g_phase[i] += t_ - t_mark;                                     \
            t_mark = t_;                                                   \
        }                                                                  \
    } while (0)
#define ECGB_MARK_DECL unsigned long long t_mark = now_ns()
#define ECGB_MARK_RESET t_mark = now_ns()
#else
#define ECGB_MARK(i) do { } while (0)
#define ECGB_MARK_DECL do { } while (0)
#define ECGB_MARK_RESET do { } while (0)
#endif

// ------------------------------------------------------------------ init / count

__global__ void bytes_to_tokens_kernel(const uint8_t *__restrict__ text, uint16_t *__restrict__ tok, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) tok[i] = text[i];
}

// get_stats (lib.rs:28-48): all overlapping windows (i, i+1), i in [0, n-1), plus the
// window that straddles into the next shard when right_tok is a token.  Block-private
// shared-memory hash first (the initial corpus has few distinct, very hot pairs).
constexpr int kCountSlots = 4096;
__global__ void __launch_bounds__(256) count_kernel(const uint16_t *__restrict__ tok, uint64_t n, uint32_t right_tok,
                                                    PairTable out) {
    __shared__ uint32_t s_keys[kCountSlots];
    __shared__ uint32_t s_cnt[kCountSlots];
    for (int i = threadIdx.x; i < kCountSlots; i += blockDim.x) { s_keys[i] = kEmptyKey; s_cnt[i] = 0; }
    __syncthreads();
    const uint64_t nwin = n == 0 ? 0 : (n - 1) + (right_tok != kSentinel ? 1 : 0);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (nwin + stride - 1) / stride;
    for (uint64_t it = 0; it < rounds; it++) {
        const uint64_t i = it * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = i < nwin;
        uint32_t key = kEmptyKey;
        if (valid) {
            const uint32_t l = tok[i];
            const uint32_t r = (i + 1 < n) ? (uint32_t)tok[i + 1] : right_tok;
            key = mk(l, r);
        }
        const unsigned m = __match_any_sync(0xffffffffu, key);
        if (valid && (threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) {
            const uint32_t c = __popc(m);
            uint32_t slot = hash_key(key) & (kCountSlots - 1);
            bool placed = false;
            for (int probes = 0; probes < 64; probes++) {
                uint32_t k = s_keys[slot];
                if (k == kEmptyKey) {
                    uint32_t old = atomicCAS(&s_keys[slot], kEmptyKey, key);
                    k = old == kEmptyKey ? key : old;
                }
                if (k == key) { atomicAdd(&s_cnt[slot], c); placed = true; break; }
                slot = (slot + 1) & (kCountSlots - 1);
            }
            if (!placed) table_add(out, key, (long long)c);
        }
        // a block-private counter is 32 bit: flush long before it can wrap
        if ((it & 0xFFFFF) == 0xFFFFF) {
            __syncthreads();
            for (int s = threadIdx.x; s < kCountSlots; s += blockDim.x)
                if (s_keys[s] != kEmptyKey && s_cnt[s]) { table_add(out, s_keys[s], (long long)s_cnt[s]); s_cnt[s] = 0; }
            __syncthreads();
        }
    }
    __syncthreads();
    for (int s = threadIdx.x; s < kCountSlots; s += blockDim.x)
        if (s_keys[s] != kEmptyKey && s_cnt[s]) table_add(out, s_keys[s], (long long)s_cnt[s]);
}

// ------------------------------------------------------------------ argmax

__device__ __forceinline__ Best better(const Best &x, const Best &y) {
    if (x.count != y.count) return x.count > y.count ? x : y;
    if (x.count == 0) return x;
    Best r = x.key < y.key ? x : y;
    r.ntied = x.ntied + y.ntied;
    return r;
}

__device__ __forceinline__ Best warp_best(Best v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best w;
        w.count = __shfl_xor_sync(0xffffffffu, v.count, o);
        w.key = __shfl_xor_sync(0xffffffffu, v.key, o);
        w.ntied = __shfl_xor_sync(0xffffffffu, v.ntied, o);
        v = better(v, w);
    }
    return v;
}

__device__ Best block_best(Best v) {
    __shared__ Best s_w[32];
    v = warp_best(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) s_w[warp] = v;
    __syncthreads();
    Best r{0, kEmptyKey, 0};
    if (warp == 0) {
        if (lane < nw) r = s_w[lane];
        r = warp_best(r);
    }
    return r;  // valid in warp 0
}

// lib.rs:92-94 with the deterministic tie rule.  The last block to finish folds the
// partials into best[step] and prepares this rank's boundary record for that pair.
__global__ void __launch_bounds__(256) argmax_kernel(TrainView v, uint32_t step) {
    const PairTable &t = v.main;
    Best mine{0, kEmptyKey, 0};
    const uint64_t cap = (uint64_t)t.mask + 1;
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t k = t.keys[s];
        if (k == kEmptyKey) continue;
        const unsigned long long c = t.cnt[s];
        if (c == 0) continue;
        mine = better(mine, Best{c, k, 1});
    }
    Best b = block_best(mine);
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        v.partial[blockIdx.x] = b;
        __threadfence();
        const uint32_t done = atomicAdd(&v.dev->argmax_done, 1u);
        s_last = done == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    Best p{0, kEmptyKey, 0};
    for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) p = better(p, v.partial[i]);
    Best fin = block_best(p);
    if (threadIdx.x == 0) {
        v.best[step] = fin;
        v.dev->argmax_done = 0;
        if (fin.count == 0) atomicMin(&v.dev->done_step, step);
        // boundary record of this shard for the winning pair
        const uint16_t *tok = v.tok[step & 1];
        const unsigned long long n = v.dev->n[step & 1];
        Boundary bd;
        memset(&bd, 0, sizeof(bd));
        bd.n_lo = (uint32_t)n;
        bd.n_hi = (uint32_t)(n >> 32);
        for (int i = 0; i < 3; i++) bd.first[i] = (unsigned long long)i < n ? (uint32_t)tok[i] : kSentinel;
        bd.last[1] = n >= 1 ? (uint32_t)tok[n - 1] : kSentinel;
        bd.last[0] = n >= 2 ? (uint32_t)tok[n - 2] : kSentinel;
        const uint32_t a = fin.key >> 16, bb = fin.key & 0xFFFFu;
        if (fin.count != 0 && a == bb && v.world > 1) {
            unsigned long long run = 0;
            while (run < n && tok[n - 1 - run] == a) run++;
            bd.trail_par = (uint32_t)(run & 1);
            bd.all_a = run == n ? 1u : 0u;
        }
        *v.boundary = bd;
    }
}

// ------------------------------------------------------------------ merge

// tile status word: [63:62] flag (1 = aggregate, 2 = inclusive prefix) | [61:42] step + 1 | [41:0] count
constexpr unsigned long long kCountMask = (1ull << 42) - 1;
__device__ __forceinline__ unsigned long long pack_status(uint32_t flag, uint32_t step, unsigned long long count) {
    return ((unsigned long long)flag << 62) | ((unsigned long long)((step + 1) & 0xFFFFF) << 42) | (count & kCountMask);
}

// rec(r) = boundary record of shard r
template <class Rec>
__device__ __forceinline__ Halo make_halo_from(Rec rec, bool have, int rank, int world, uint32_t a, uint32_t b) {
    Halo h;
    h.nl = h.nr = 0;
    h.L[0] = h.L[1] = kSentinel;
    h.R[0] = h.R[1] = h.R[2] = kSentinel;
    h.par_in = 0;
    if (world <= 1 || !have) return h;
    // left context: the last two tokens of the stream before this shard
    uint32_t got[2];
    int ng = 0;
    for (int r = rank - 1; r >= 0 && ng < 2; r--) {
        const unsigned long long n = ((unsigned long long)rec(r).n_hi << 32) | rec(r).n_lo;
        if (n >= 1) got[ng++] = rec(r).last[1];
        if (n >= 2 && ng < 2) got[ng++] = rec(r).last[0];
    }
    h.nl = ng;
    if (ng >= 1) h.L[1] = got[0];
    if (ng >= 2) h.L[0] = got[1];
    // right context: the first three tokens of the stream after this shard
    int nr = 0;
    for (int r = rank + 1; r < world && nr < 3; r++) {
        const unsigned long long n = ((unsigned long long)rec(r).n_hi << 32) | rec(r).n_lo;
        for (int i = 0; i < 3 && (unsigned long long)i < n && nr < 3; i++) h.R[nr++] = rec(r).first[i];
    }
    h.nr = nr;
    if (a == b) {
        uint32_t par = 0;
        for (int r = rank - 1; r >= 0; r--) {
            const unsigned long long n = ((unsigned long long)rec(r).n_hi << 32) | rec(r).n_lo;
            if (n == 0) continue;
            if (rec(r).all_a) { par ^= (uint32_t)(n & 1); continue; }
            par ^= rec(r).trail_par;
            break;
        }
        h.par_in = par;
    }
    return h;
}

__device__ Halo make_halo(const Boundary *all, int rank, int world, uint32_t a, uint32_t b) {
    return make_halo_from([&](int r) -> const Boundary & { return all[r]; }, all != nullptr, rank, world, a, b);
}

struct MergeSmem {
    Halo halo;
    long long tile;
    int scan[kTPB / 32];
    long long lastnon[kTPB / 32];
    unsigned long long prefix;
    long long tile_lastnon;
    unsigned long long tau_val;  // candidate threshold of the table the patches go to, read once per pass
    int chunk_n;               // resident tail: tokens of this CTA's chunk
    uint32_t carry_ctx[2];     // resident tail: last two input tokens of the previous tile
    Boundary bd_near[5];       // resident tail: records of chunks blockIdx-2 .. blockIdx+2
    Boundary bd_peer[kMaxWorld];  // persistent sharded loop: every rank's shard record for this step
    PeerPush push;             // persistent sharded loop: where this step's patches go
    uint32_t peer_cnt[kMaxWorld];  // persistent sharded loop: entries of each peer's patch list of this step
    __align__(8) unsigned long long mbar[kChunkTiles];  // chunks in global memory: one mbarrier per slot of the tile ring
    // in[8 + q] = token at tile position q; in[6..7] / in[8 + kTile ..] = 2 / 3 tokens of context
    __align__(16) uint16_t in[kTile + 16];
    // kept tokens of the tile, compacted; 32-bit words XOR-swizzled (see stage_index)
    __align__(16) uint32_t out[kTile / 2 + 32];
    PatchTable patch;
};

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// Resident tail: boundary record of this CTA's chunk for the pair (a, b) (the same record a rank
// publishes for its shard).  Warp 0 only.
__device__ __forceinline__ void chunk_boundary(const uint16_t *chunk, int n, uint32_t a, bool same, Boundary *dst) {
    const int lane = threadIdx.x & 31;
    int run = 0;
    if (same) {  // length of the trailing run of a
        int found = -1;
        bool hit = false;
        for (int p = n - 1; p >= 0 && !hit; p -= 32) {
            const int q = p - lane;
            const bool nonx = q >= 0 && chunk[q] != a;
            const unsigned m = __ballot_sync(0xffffffffu, nonx);
            if (m) { found = p - (__ffs(m) - 1); hit = true; }
        }
        run = n - 1 - found;
    }
    if (lane == 0) {
        Boundary bd;
        memset(&bd, 0, sizeof(bd));
        bd.n_lo = (uint32_t)n;
        for (int i = 0; i < 3; i++) bd.first[i] = i < n ? (uint32_t)chunk[i] : kSentinel;
        bd.last[1] = n >= 1 ? (uint32_t)chunk[n - 1] : kSentinel;
        bd.last[0] = n >= 2 ? (uint32_t)chunk[n - 2] : kSentinel;
        bd.trail_par = (uint32_t)(run & 1);
        bd.all_a = same && run == n ? 1u : 0u;
        *dst = bd;
    }
}

// Staging swizzle.  A thread writes ~16 consecutive tokens, so the lanes of a warp start
// 8 words apart and would pile onto 4 banks; XOR-ing the low 3 bits of the word index with
// the low 3 bits of its 32-word row spreads them over all 32 banks, and a warp reading 32
// consecutive words still touches every bank at most twice.
__device__ __forceinline__ uint32_t stage_word(uint32_t w) { return w ^ ((w >> 5) & 7u); }
__device__ __forceinline__ uint32_t stage_index(uint32_t x) { return x ^ ((x >> 5) & 14u); }  // 16-bit index

// merge (lib.rs:10-26) + incremental get_stats.  `upd` receives the histogram patches.
// TICKETS: tiles are handed out by an atomic counter (any grid size); otherwise tile =
// blockIdx + k * gridDim, which needs every block to be co-resident (cooperative launch).
// Spin until the 8-byte unit at p (this rank's area, written by a peer over NVLink) carries `tag`; returns its
// payload.  Gives up -- returning 0 with *abort set -- after timeout_ns or when another wait already gave up.
__device__ __forceinline__ uint32_t wait_unit(const void *p, uint32_t tag, volatile unsigned int *abort,
                                              unsigned long long timeout_ns, uint32_t where) {
    unsigned long long t0 = 0;
    for (unsigned int spins = 0;; spins++) {
        unsigned long long u;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(u) : "l"(p) : "memory");
        if ((uint32_t)(u >> 32) == tag) return (uint32_t)u;
        if ((spins & 255u) == 255u) {
            if (*abort) return 0u;
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > timeout_ns) { *abort = where | 0x80000000u; return 0u; }  // what was being waited for
        }
    }
}
__device__ __forceinline__ void st_unit(void *p, uint32_t payload, uint32_t tag) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(peer_unit(payload, tag)) : "memory");
}

// ---- tile ring of a chunk that lives in global memory: the bulk-copy engine (cp.async.bulk, "TMA" 1-D) streams
// the CTA's next tiles into shared memory while the current one is merged; an mbarrier per slot counts the bytes.
// Built only with -DECGB_TILE_RING: measured on B200 (1.2e9-symbol corpus, 2 000 merges) it is bit-exact but
// SLOWER than the plain 128-bit loads (0.677 s vs 0.613 s) -- the pass is bound by its barriers and instruction
// stream, not by load latency (ncu: barrier stalls 4.6 of 10 warp-cycles, issue 55 %), so prefetching buys nothing.
struct TileRing {
    uint16_t *slots;        // kChunkTiles x kTile tokens (the dynamic shared memory the resident chunk uses later)
    uint32_t uses[kChunkTiles];  // copies issued into each slot so far (parity of the phase to wait for); same in every thread
};
__device__ __forceinline__ void ring_init(MergeSmem &sm) {  // one thread, once per kernel
#pragma unroll
    for (int s = 0; s < kChunkTiles; s++)
        asm volatile("mbarrier.init.shared.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&sm.mbar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void ring_issue(MergeSmem &sm, const TileRing &ring, int slot, const uint16_t *src, uint32_t bytes) {  // one thread
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&sm.mbar[slot]);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring.slots + (size_t)slot * kTile);
    asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void ring_wait(MergeSmem &sm, int slot, uint32_t parity) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&sm.mbar[slot]);
    uint32_t ok = 0;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}

template <bool TICKETS, bool RESIDENT, bool RING = false>
__device__ __forceinline__ void merge_pass(const TrainView &v, uint32_t step, const Best bb,
                                           const Boundary *__restrict__ all_bd, const PairTable &upd, MergeSmem &sm,
                                           uint16_t *chunk, const PeerView *pv = nullptr, TileRing *ring = nullptr) {
    static_assert(!RING || RESIDENT, "the tile ring feeds chunks");
    // Resident tail (chunk != nullptr, cooperative kernel only): the stream lives in the CTAs' shared
    // memory, one contiguous chunk each, and is merged in place; the chunks are shards exactly like the
    // ranks of a sharded run (all_bd = every CTA's boundary record), so no output offsets are needed.
    static_assert(!(TICKETS && RESIDENT), "the resident tail belongs to the cooperative kernel");
    constexpr bool resident = RESIDENT;
    const uint32_t a = bb.key >> 16, b = bb.key & 0xFFFFu, z = v.new_ids != nullptr ? v.new_ids[step] : 256u + step;
    const uint16_t *__restrict__ in = resident ? chunk : v.tok[step & 1];
    uint16_t *__restrict__ out = resident ? chunk : v.tok[(step + 1) & 1];
    const long long n = resident ? (long long)sm.chunk_n : (long long)v.dev->n[step & 1];
    const long long ntiles = n == 0 ? 1 : (n + kTile - 1) / kTile;
    const bool same = a == b;
    const uint32_t ab = a | (b << 16);  // a site, as the 32-bit word of two adjacent tokens

    ECGB_MARK_DECL;
    __syncthreads();
    if (threadIdx.x == 32) sm.tau_val = upd.tau != nullptr ? *upd.tau : ~0ull;
    const bool sharded = pv != nullptr && pv->world > 1;  // persistent sharded loop: the peers' shard records live in this rank's area
    if (sharded) {
        const int par = (int)(step & 1u);
        const uint8_t *mine = pv->area[pv->rank];
        const uint32_t tag = peer_tag(pv->epoch, step);
        if (threadIdx.x < (unsigned)pv->world * 8u) {
            // the peers' shard records for this step's stream, one thread per unit (the run fields arrive in the
            // second exchange of an (x,x) step)
            const int r = (int)(threadIdx.x >> 3), u = (int)(threadIdx.x & 7u);
            if (r != pv->rank && (u < 5 || (u == 5 && same))) {
                const uint32_t val = wait_unit(mine + peer_rec_off(pv->world, par, r) + 8 * u, tag, v.abort, pv->timeout_ns,
                                               (1u << 28) | ((uint32_t)u << 24) | ((uint32_t)r << 20) | (step & 0xFFFFFu));
                Boundary &bd = sm.bd_peer[r];
                if (u == 0) bd.n_lo = val;
                else if (u == 1) bd.n_hi = val;
                else if (u == 2) { bd.first[0] = val & 0xFFFFu; bd.first[1] = val >> 16; }
                else if (u == 3) { bd.first[2] = val & 0xFFFFu; bd.last[0] = val >> 16; }
                else if (u == 4) bd.last[1] = val;
                else { bd.trail_par = val & 1u; bd.all_a = (val >> 1) & 1u; }
            }
        }
        if (threadIdx.x == kTPB - 1) {  // this step's patches: local table + every peer's inbox
            PeerPush &pp = sm.push;
            pp.out_count = pv->out_count + par;
            pp.overflow = pv->overflow;
            pp.cap = pv->cap;
            pp.tag = tag;
            pp.rank = pv->rank;
            pp.world = pv->world;
            for (int r = 0; r < pv->world; r++)
                pp.dst[r] = reinterpret_cast<uint4 *>(pv->area[r] + peer_ent_off(pv->world, pv->cap, par, pv->rank));
        }
    }
    if (resident) {
        // the neighbours' records in one round trip; make_halo rarely needs anything further away
        const int me = (int)blockIdx.x;
        if (threadIdx.x >= 64 && threadIdx.x < 64 + 5 * kBoundaryWords) {
            const int q = (threadIdx.x - 64) / kBoundaryWords, wd = (threadIdx.x - 64) % kBoundaryWords;
            const int r = me - 2 + q;
            if (r >= 0 && r < (int)gridDim.x)
                reinterpret_cast<uint32_t *>(&sm.bd_near[q])[wd] = reinterpret_cast<const uint32_t *>(&all_bd[r])[wd];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            auto local = [&](int r) -> const Boundary & { return (r >= me - 2 && r <= me + 2) ? sm.bd_near[r - me + 2] : all_bd[r]; };
            if (sharded) {
                // one chain of shards: the ranks before this one (one record each), this rank's chunks, the ranks after
                const int rk = pv->rank, G = (int)gridDim.x;
                sm.halo = make_halo_from(
                    [&](int i) -> const Boundary & { return i < rk ? sm.bd_peer[i] : (i < rk + G ? local(i - rk) : sm.bd_peer[i - G + 1]); },
                    true, rk + me, pv->world - 1 + G, a, b);
            } else {
                sm.halo = make_halo_from(local, true, me, (int)gridDim.x, a, b);
            }
        }
    } else if (sharded) {
        __syncthreads();
        if (threadIdx.x == 0) sm.halo = make_halo(sm.bd_peer, pv->rank, pv->world, a, b);
    } else if (threadIdx.x == 0) {
        sm.halo = make_halo(all_bd, v.rank, v.world, a, b);
    }
    patch_clear(sm.patch);
    if (threadIdx.x == 0) {
        sm.patch.gate = TICKETS ? nullptr : v.arrive;
        sm.patch.gate_target = (step + 1u) * gridDim.x;
    }
    __syncthreads();
    ECGB_MARK(1);
    PairTable updp = upd;  // the view of the table the patches go through
    if (sharded) updp.push = &sm.push;
    const Halo h = sm.halo;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint16_t *const stage16 = reinterpret_cast<uint16_t *>(sm.out);

    // token at shard-local position p, including halo context
    auto tok_at = [&](long long p) -> uint32_t {
        if (p >= 0 && p < n) return in[p];
        if (p < 0) return (-p <= (long long)h.nl) ? h.L[2 + p] : kSentinel;
        const long long q = p - n;
        return q < (long long)h.nr ? h.R[q] : kSentinel;
    };

    // RING: tile t of the chunk goes through slot t % kChunkTiles; two tiles are in flight ahead of the merge
    auto tile_bytes = [&](long long t) -> uint32_t {
        const long long valid = min((long long)kTile, n - t * kTile);
        return (uint32_t)((valid * 2 + 15) & ~15ll);  // the token buffers end with slack, chunk starts are 16-byte aligned
    };
    uint32_t ring_used[kChunkTiles] = {0, 0, 0};  // copies issued in this pass, per slot
    if (RING) {
        if (threadIdx.x == 0 && n > 0) {
            ring_issue(sm, *ring, 0, in, tile_bytes(0));
            if (ntiles > 1) ring_issue(sm, *ring, 1, in + kTile, tile_bytes(1));
        }
        if (n > 0) { ring_used[0]++; if (ntiles > 1) ring_used[1]++; }
    }

    unsigned long long chunk_fill = 0;  // resident: tokens of the chunk written so far
    for (long long round = 0;; round++) {
        long long tile;
        if (TICKETS) {
            __syncthreads();
            if (threadIdx.x == 0) sm.tile = (long long)atomicAdd(&v.tickets[step], 1u);
            __syncthreads();
            tile = sm.tile;
        } else {
            tile = resident ? round : (long long)blockIdx.x + round * (long long)gridDim.x;
            // Chunks: no barrier between tiles.  What the next tile overwrites first is sm.in, which the previous tile
            // stopped reading before its last barrier (the write-out reads sm.out only); sm.out is written again two
            // barriers into the next tile; the in-place write-out ends where the next tile's loads begin.
            if (!resident) __syncthreads();  // shared staging of the previous tile is free again
        }
        if (tile >= ntiles) break;
        const long long tbase = tile * kTile;
        const long long base = tbase + (long long)threadIdx.x * kIPT;
        const int tile_valid = (int)min((long long)kTile, n - tbase);               // tokens of the stream in this tile
        const int nvalid = min(kIPT, max(0, tile_valid - (int)threadIdx.x * kIPT));  // ... in this thread's range
        const uint32_t validmask = (1u << nvalid) - 1u;

        const uint16_t *next_slot = nullptr;  // RING: the tile after this one, already in shared memory
        if (RING && n > 0) {
            const int slot = (int)(tile % kChunkTiles);
            // the slot of tile + 2 held tile - 1, which every thread finished reading before the barrier above
            if (tile + 2 < ntiles) {
                const int s2 = (int)((tile + 2) % kChunkTiles);
                if (threadIdx.x == 0) ring_issue(sm, *ring, s2, in + (tile + 2) * kTile, tile_bytes(tile + 2));
                ring_used[s2]++;
            }
            ring_wait(sm, slot, (ring->uses[slot] + ring_used[slot] - 1u) & 1u);
            if (tile + 1 < ntiles) {
                const int s1 = (int)((tile + 1) % kChunkTiles);
                ring_wait(sm, s1, (ring->uses[s1] + ring_used[s1] - 1u) & 1u);
                next_slot = ring->slots + (size_t)s1 * kTile;
            }
        }
        // ---- 16 tokens per thread, two per 32-bit word (token 2j = low half of w[j]) ----
        uint32_t w[kIPT / 2];
        if (nvalid == kIPT) {
            const uint16_t *src16 = RING ? ring->slots + (size_t)(tile % kChunkTiles) * kTile + threadIdx.x * kIPT : in + base;
            const uint4 v0 = *reinterpret_cast<const uint4 *>(src16);
            const uint4 v1 = *reinterpret_cast<const uint4 *>(src16 + 8);
            w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w;
            w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
        } else {  // end of the shard: positions >= n show the right halo (or the sentinel)
#pragma unroll
            for (int j = 0; j < kIPT / 2; j++) w[j] = tok_at(base + 2 * j) | (tok_at(base + 2 * j + 1) << 16);
        }
        {
            uint4 *dst = reinterpret_cast<uint4 *>(&sm.in[8 + threadIdx.x * kIPT]);
            dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
        // (resident: the tokens before this tile have already been merged in place; the previous tile left a copy)
        if (threadIdx.x < 2)
            sm.in[6 + threadIdx.x] = (uint16_t)((resident && tile > 0) ? sm.carry_ctx[threadIdx.x] : tok_at(tbase - 2 + threadIdx.x));
        if (threadIdx.x >= 2 && threadIdx.x < 5) {
            const long long pr = tbase + kTile + threadIdx.x - 2;  // right context: the first tokens of the next tile
            sm.in[8 + kTile + threadIdx.x - 2] = (uint16_t)((RING && next_slot != nullptr && pr < n) ? next_slot[threadIdx.x - 2] : tok_at(pr));
        }
        __syncthreads();
        ECGB_MARK(2);
        if (resident && threadIdx.x == kTPB - 1) { sm.carry_ctx[0] = w[7] & 0xFFFFu; sm.carry_ctx[1] = w[7] >> 16; }
        // neighbours of the range: from the adjacent lanes, across warps from the shared tile
        uint32_t tprev = __shfl_up_sync(0xffffffffu, w[7] >> 16, 1);
        uint32_t tnext = __shfl_down_sync(0xffffffffu, w[0] & 0xFFFFu, 1);
        if (lane == 0) tprev = sm.in[8 + threadIdx.x * kIPT - 1];
        if (lane == 31) tnext = sm.in[8 + threadIdx.x * kIPT + kIPT];

        uint32_t site = 0, removed = 0;  // bit i = position base + i
        if (same) {
            // ---- (x,x): run offset parity needs the position of the last non-x before each token ----
            uint32_t isa = 0;  // bit i: token i is x
#pragma unroll
            for (int j = 0; j < kIPT / 2; j++) {
                if ((w[j] & 0xFFFFu) == a) isa |= 1u << (2 * j);
                if ((w[j] >> 16) == a) isa |= 2u << (2 * j);
            }
            const uint32_t non = ~isa & validmask;
            long long x = non ? base + (31 - __clz(non)) : -(1ll << 62);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {  // inclusive max-scan inside the warp
                const long long y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x = max(x, y);
            }
            if (lane == 31) sm.lastnon[warp] = x;
            if (resident) {  // tiles are taken in order: the previous tile left the position in tile_lastnon
                if (threadIdx.x == 0 && tile == 0) sm.tile_lastnon = -1 - (long long)h.par_in;
            } else if (warp == 0) {  // run entering the tile: scan backwards from the tile start
                long long found = -(1ll << 62);
                bool hit = false;
                for (long long p = tbase - 1; p >= 0 && !hit; p -= 32) {
                    const long long q = p - lane;
                    const bool nonx = q >= 0 && in[q] != a;
                    const unsigned m = __ballot_sync(0xffffffffu, nonx);
                    if (m) { found = p - (__ffs(m) - 1); hit = true; }
                }
                // reached the shard start inside the run: continue it virtually by par_in elements
                if (!hit) found = -1 - (long long)h.par_in;
                if (lane == 0) sm.tile_lastnon = found;
            }
            __syncthreads();
            long long before = sm.tile_lastnon;
            for (int q = 0; q < warp; q++) before = max(before, sm.lastnon[q]);
            const long long xe = __shfl_up_sync(0xffffffffu, x, 1);
            if (lane > 0) before = max(before, xe);
            // odd = parity of the offset inside the run of x that position i belongs to
            uint32_t odd = (uint32_t)(base - (before + 1)) & 1u;
            const uint32_t isa_next = (isa >> 1) | ((tnext == a ? 1u : 0u) << (kIPT - 1));
#pragma unroll
            for (int i = 0; i < kIPT; i++) {
                const uint32_t bit = 1u << i;
                if (isa & bit) {
                    if (odd) removed |= bit; else if (isa_next & bit) site |= bit;
                    odd ^= 1u;
                } else {
                    odd = 0;
                }
            }
            site &= validmask;
            removed &= validmask;
        } else {
            // ---- a != b: a site is the 32-bit word (a | b << 16) at an even or an odd token offset ----
#pragma unroll
            for (int j = 0; j < kIPT / 2; j++) {
                if (w[j] == ab) site |= 1u << (2 * j);
                const uint32_t hi = j + 1 < kIPT / 2 ? w[(j + 1) & 7] : tnext;
                if (__funnelshift_r(w[j], hi, 16) == ab) site |= 2u << (2 * j);
            }
            site &= validmask;
            const uint32_t before = (tprev | (w[0] << 16)) == ab ? 1u : 0u;  // a site at base - 1
            removed = ((site << 1) | before) & validmask;
        }

        // ---- histogram patches, one site at a time (see oracle/ecgb_oracle.c ecgo_train_fast) ----
        if (__any_sync(0xffffffffu, site != 0)) {
            int ns = __popc(site);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ns += __shfl_xor_sync(0xffffffffu, ns, o);
            if (lane == 0) patch_add(sm.patch, updp, mk(a, b), -ns);  // the pair itself, per warp
            const uint16_t *ctx = &sm.in[6 + threadIdx.x * kIPT];  // ctx[2 + i] = token at base + i
            uint32_t rem = site;
            while (rem) {
                const int i = __ffs(rem) - 1;
                rem &= rem - 1;
                const uint32_t tm2 = ctx[i], tm1 = ctx[1 + i], tp2 = ctx[4 + i], tp3 = ctx[5 + i];
                if (tm1 != kSentinel) {  // there is a left neighbour
                    const bool prev_site = tm2 == a && tm1 == b;  // site at p-2 (parity is implied)
                    patch_add(sm.patch, updp, mk(tm1, a), -1);
                    patch_add(sm.patch, updp, mk(prev_site ? z : tm1, z), +1);
                }
                if (tp2 != kSentinel && !(tp2 == a && tp3 == b)) {  // right neighbour that starts no site
                    patch_add(sm.patch, updp, mk(b, tp2), -1);
                    patch_add(sm.patch, updp, mk(z, tp2), +1);
                }
            }
        }

        ECGB_MARK(3);
        // ---- compaction: kept tokens -> shared staging -> coalesced stores ----
        const uint32_t keepmask = validmask & ~removed;
        const int kept = __popc(keepmask);
        int incl = kept;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) sm.scan[warp] = incl;
        __syncthreads();
        ECGB_MARK(4);
        if (resident && same && threadIdx.x == 0) {  // every thread has read tile_lastnon: carry it to the next tile
            long long m = sm.tile_lastnon;
            for (int q = 0; q < kTPB / 32; q++) m = max(m, sm.lastnon[q]);
            sm.tile_lastnon = m;
        }
        int warp_off = 0, tile_total = 0;
#pragma unroll
        for (int q = 0; q < kTPB / 32; q++) {
            const int c = sm.scan[q];
            if (q < warp) warp_off += c;
            tile_total += c;
        }
        uint32_t o = (uint32_t)(warp_off + incl - kept);
        if ((site | removed) == 0 && nvalid == kIPT && (o & 1u) == 0) {
            // untouched range at an even offset: eight whole words
#pragma unroll
            for (int j = 0; j < kIPT / 2; j++) sm.out[stage_word((o >> 1) + j)] = w[j];
        } else {
#pragma unroll
            for (int i = 0; i < kIPT; i++) {
                const uint32_t t = (i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xFFFFu);
                if ((keepmask >> i) & 1u) stage16[stage_index(o++)] = (uint16_t)(((site >> i) & 1u) ? z : t);
            }
        }

        ECGB_MARK(5);
        // decoupled look-back over the tiles that precede this one, 32 tiles per probe (warp 0)
        if (warp == 0 && !resident) {
            unsigned long long excl = 0;
            if (tile > 0) {
                if (lane == 0) atomicExch(&v.tile_status[tile], pack_status(1, step, (unsigned long long)tile_total));
                long long j = tile - 1;  // lane l looks at tile j - l
                unsigned int lb_spins = 0;
                for (;;) {
                    const long long idx = j - lane;
                    uint32_t flag = 2;            // tiles before the first one: an inclusive prefix of 0
                    unsigned long long cntv = 0;
                    if (idx >= 0) {
                        const unsigned long long sdw = ld_status(&v.tile_status[idx]);
                        const uint32_t ep = (uint32_t)((sdw >> 42) & 0xFFFFF);
                        flag = ep == ((step + 1) & 0xFFFFF) ? (uint32_t)(sdw >> 62) : 0u;
                        cntv = sdw & kCountMask;
                    }
                    const unsigned notready = __ballot_sync(0xffffffffu, flag == 0);
                    const unsigned inclm = __ballot_sync(0xffffffffu, flag == 2);
                    const int first_nr = notready ? __ffs(notready) - 1 : 32;
                    const int first_in = inclm ? __ffs(inclm) - 1 : 32;
                    if (first_in < first_nr) {  // everything up to an inclusive prefix is published
                        unsigned long long c = lane <= first_in ? cntv : 0ull;
#pragma unroll
                        for (int o2 = 16; o2 > 0; o2 >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o2);
                        excl += c;
                        break;
                    }
                    if (first_nr == 32) {  // 32 aggregates: take them all and look further back
                        unsigned long long c = cntv;
#pragma unroll
                        for (int o2 = 16; o2 > 0; o2 >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o2);
                        excl += c;
                        j -= 32;
                    }
                    // else: a needed tile is not published yet -> probe again
                    if ((++lb_spins & 1023u) == 0 && v.abort != nullptr && *reinterpret_cast<volatile unsigned int *>(v.abort)) break;
                }
            }
            if (lane == 0) {
                __threadfence();
                atomicExch(&v.tile_status[tile], pack_status(2, step, excl + (unsigned long long)tile_total));
                sm.prefix = excl;
                if (tile == ntiles - 1) {
                    v.dev->n[(step + 1) & 1] = excl + (unsigned long long)tile_total;
                    v.n_hist[step + 1] = excl + (unsigned long long)tile_total;
                }
            }
        }
        ECGB_MARK(6);
        __syncthreads();
        ECGB_MARK(7);
        const unsigned long long gofs = resident ? chunk_fill : sm.prefix;  // resident: the chunk is compacted in place
        chunk_fill += (unsigned long long)tile_total;
        // ---- write-out as 32-bit words; an odd output offset shifts the word boundary by one token ----
        const int lead = (int)(gofs & 1ull) & (tile_total > 0 ? 1 : 0);
        const int nwords = (tile_total - lead) >> 1;
        uint32_t *__restrict__ out32 = reinterpret_cast<uint32_t *>(out + gofs + lead);
        if (lead) {
            for (int m = threadIdx.x; m < nwords; m += kTPB) {  // word m = staged tokens 2m+1, 2m+2
                const uint32_t lo = sm.out[stage_word((uint32_t)m)], hi = sm.out[stage_word((uint32_t)m + 1u)];
                out32[m] = __funnelshift_r(lo, hi, 16);
            }
            if (threadIdx.x == 0) out[gofs] = stage16[0];
        } else {
            for (int m = threadIdx.x; m < nwords; m += kTPB) out32[m] = sm.out[stage_word((uint32_t)m)];
        }
        if (threadIdx.x == 32 && ((tile_total - lead) & 1))
            out[gofs + tile_total - 1] = stage16[stage_index((uint32_t)tile_total - 1u)];
        ECGB_MARK(8);
    }
    if (RING) {
#pragma unroll
        for (int q = 0; q < kChunkTiles; q++) ring->uses[q] += ring_used[q];
        // the next pass reads this chunk through the bulk-copy engine (async proxy): order this thread's stores before it
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    if (resident) {
        __syncthreads();  // the chunk is complete
        // record for the next step (other parity: slower CTAs may still be reading this step's records);
        // an (x,x) step adds the run information and a grid barrier of its own
        if (warp == 0) chunk_boundary(chunk, (int)chunk_fill, kSentinel, false, &v.cta_bd[((step + 1) & 1) * v.cta_stride + blockIdx.x]);
        if (threadIdx.x == 0) {
            sm.chunk_n = (int)chunk_fill;
            if (chunk_fill) atomicAdd(&v.n_hist[step + 1], chunk_fill);  // stream length after this step (zero-initialised)
        }
    }
    if (!TICKETS) {
        // the histogram may only change once every CTA has taken this step's argmax from it (CTAs scan
        // short candidate lists on their own, without a grid barrier): cumulative arrival counter
        if (threadIdx.x == 0) {
            const unsigned int target = (step + 1u) * gridDim.x;
            unsigned int spins = 0;
            while (*reinterpret_cast<volatile unsigned int *>(v.arrive) < target) {
                if ((++spins & 1023u) == 0 && *reinterpret_cast<volatile unsigned int *>(v.abort)) break;
            }
        }
    }
    __syncthreads();
    patch_flush(sm.patch, updp, sm.tau_val);
    ECGB_MARK(9);
}

__global__ void __launch_bounds__(kTPB) merge_kernel(TrainView v, uint32_t step, const Boundary *__restrict__ all_bd,
                                                     PairTable upd) {
    __shared__ MergeSmem sm;
    if (step == kStepFromDevice) step = v.dev->cur_step;
    if (step >= v.dev->max_steps) return;
    const Best bb = v.best[step];
    if (bb.count == 0) return;  // no pair left (lib.rs:88-90); done_step was recorded by argmax_kernel
    merge_pass<true, false>(v, step, bb, all_bd, upd, sm, nullptr);
}

// The whole single-device training loop (lib.rs:85-117) as ONE persistent cooperative
// kernel: per step, a grid-wide argmax over the pair table, a grid barrier, the streaming
// merge pass, a grid barrier.  No launches and no host round trips inside the loop.
// grid-wide fold of per-block partial maxima; every block ends up with the same result
// Grid barrier of the persistent sharded loop: cumulative arrival counter, and -- unlike cg::grid_group::sync --
// a way out: every spin loop of that kernel also watches *abort, which a timed-out wait for a peer sets, so a
// rank whose peer died ends with an error instead of hanging the GPU.
struct AbortableGrid {
    unsigned int *ctr;
    volatile unsigned int *abort;
    unsigned int gen;
    __device__ __forceinline__ void sync() {
        __syncthreads();
        gen++;
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(ctr, 1u);
            const unsigned int target = gen * gridDim.x;
            unsigned int spins = 0;
            while ((int)(*reinterpret_cast<volatile unsigned int *>(ctr) - target) < 0) {
                if ((++spins & 255u) == 0 && *abort) break;
            }
            __threadfence();
        }
        __syncthreads();
    }
};

template <class Grid>
__device__ __forceinline__ Best grid_best(Grid &grid, const TrainView &v, Best mine, Best *s_best) {
    Best bb = block_best(mine);
    if (threadIdx.x == 0) v.partial[blockIdx.x] = bb;
    grid.sync();
    Best p{0, kEmptyKey, 0};
    for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) p = better(p, v.partial[i]);
    Best fin = block_best(p);
    __syncthreads();
    if (threadIdx.x == 0) *s_best = fin;
    __syncthreads();
    fin = *s_best;
    grid.sync();  // partial[] may be rewritten after this point
    return fin;
}

// argmax (lib.rs:92-94) for a cooperative grid; every thread returns the same winner.
template <class Grid>
__device__ __forceinline__ Best grid_argmax(Grid &grid, const TrainView &v, Best *s_best_p) {
    Best &s_best = *s_best_p;
    const PairTable &t = v.main;
    const uint64_t cap = (uint64_t)t.mask + 1;
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t gthreads = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long *tau = const_cast<unsigned long long *>(t.tau);
    // ---- argmax (lib.rs:92-94): over the candidate list; every pair whose count is >= tau
    //      is listed, so a listed maximum >= tau is the global one (with all its ties)
    const unsigned long long cur_tau = *tau;
    const uint32_t nc = *t.ncand;
    Best mine{0, kEmptyKey, 0};
    Best fin;
    if (nc <= v.redundant_max) {
        // very short list (the usual case in the long tail): every CTA scans it and reaches the same
        // winner by itself -- no grid barrier between the histogram and the merge pass
        for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x) {
            const uint32_t slot = t.cand[i];
            const unsigned long long c = t.cnt[slot];
            if (c != 0) mine = better(mine, Best{c, t.keys[slot], 1});
        }
        const Best b0 = block_best(mine);
        __syncthreads();
        if (threadIdx.x == 0) s_best = b0;
        __syncthreads();
        fin = s_best;
    } else if (nc <= 8 * kTPB) {
        // short list: block 0 scans it alone, one grid barrier publishes the winner
        if (blockIdx.x == 0) {
            for (uint32_t i = threadIdx.x; i < nc; i += blockDim.x) {
                const uint32_t slot = t.cand[i];
                const unsigned long long c = t.cnt[slot];
                if (c != 0) mine = better(mine, Best{c, t.keys[slot], 1});
            }
            const Best b0 = block_best(mine);
            if (threadIdx.x == 0) v.partial[v.cta_stride] = b0;
        }
        grid.sync();
        fin = v.partial[v.cta_stride];
    } else {
        for (uint64_t i = gtid; i < nc; i += gthreads) {
            const uint32_t slot = t.cand[i];
            const unsigned long long c = t.cnt[slot];
            if (c != 0) mine = better(mine, Best{c, t.keys[slot], 1});
        }
        fin = grid_best(grid, v, mine, &s_best);
    }
    if (fin.count == 0 || fin.count < cur_tau) {
        // the list is exhausted: full scan for the true maximum, then rebuild the list with
        // a lower threshold (counts only decay, so this happens O(log) times per run)
        mine = Best{0, kEmptyKey, 0};
        for (uint64_t s = gtid; s < cap; s += gthreads) {
            const uint32_t k = t.keys[s];
            if (k == kEmptyKey) continue;
            const unsigned long long c = t.cnt[s];
            if (c != 0) mine = better(mine, Best{c, k, 1});
        }
        fin = grid_best(grid, v, mine, &s_best);
        const unsigned long long new_tau = fin.count - fin.count / 4 > 0 ? fin.count - fin.count / 4 : 1;
        if (gtid == 0) { *tau = new_tau; *t.ncand = 0; }
        for (uint64_t w = gtid; w < (cap + 31) / 32; w += gthreads) t.inbits[w] = 0;
        grid.sync();
        for (uint64_t s = gtid; s < cap; s += gthreads) {
            if (t.keys[s] == kEmptyKey || t.cnt[s] < new_tau) continue;
            atomicOr(&t.inbits[s >> 5], 1u << (s & 31));
            t.cand[atomicAdd(t.ncand, 1u)] = (uint32_t)s;
        }
        grid.sync();
    }
    return fin;
}

// Where a CTA's piece of the token stream lives in the persistent kernels.  Default (resident_ok = 2): the
// stream is cut into one contiguous chunk per CTA before the first step and stays that way -- each chunk is
// merged in place, in the first token buffer while it is longer than kChunkCap and in the CTA's shared memory
// from then on (every CTA moves on its own: the neighbours only ever see its boundary record).  No output
// offsets, hence no look-back chain, no second buffer, and the CTAs stream independently of each other.
// resident_ok = 1: the round-1 scheme (one stream, ping-pong buffers, decoupled look-back; chunks only once
// the whole stream fits in shared memory); 0: never chunk.  Both kept for A/B runs (ECGB_RESIDENT_TAIL).
struct ChunkWhere {
    uint16_t *g = nullptr;   // the chunk's region in the first token buffer (chunked from the start)
    bool res_mode = false;   // the stream is held as chunks
    bool in_smem = false;    // this CTA's chunk is in shared memory
    bool fresh = false;      // chunks were just formed: no boundary records published yet
    bool from_start = false; // resident_ok = 2
    TileRing ring;           // tile ring of a chunk in global memory
};

__device__ __forceinline__ void chunk_step_begin(const TrainView &v, uint32_t step, MergeSmem &sm, uint16_t *chunk_smem, ChunkWhere &w) {
    if (!w.res_mode && v.resident_ok == 2u && step == 0) {
        const unsigned long long n = v.dev->n[0];
        const unsigned long long clen = (((n + gridDim.x - 1) / gridDim.x) + 7ull) & ~7ull;  // 16-byte aligned chunk starts
        const unsigned long long lo = min(n, (unsigned long long)blockIdx.x * clen);
        const unsigned long long hi = min(n, lo + clen);
        w.g = v.tok[0] + lo;
        if (threadIdx.x == 0) {
            sm.chunk_n = (int)(hi - lo);
            ring_init(sm);
        }
        w.ring.slots = chunk_smem;
        w.ring.uses[0] = w.ring.uses[1] = w.ring.uses[2] = 0;
        w.res_mode = w.fresh = w.from_start = true;
        __syncthreads();
    } else if (!w.res_mode && v.resident_ok == 1u) {
        // the stream now fits in the CTAs' shared memory: load this CTA's chunk and stay on chip
        const unsigned long long n = v.dev->n[step & 1];
        const unsigned long long clen = (((n + gridDim.x - 1) / gridDim.x) + 7ull) & ~7ull;
        if (clen <= (unsigned long long)kChunkCap) {
            const unsigned long long lo = min(n, (unsigned long long)blockIdx.x * clen);
            const unsigned long long hi = min(n, lo + clen);
            w.g = v.tok[step & 1] + lo;  // lo is a multiple of 8 tokens: 16-byte aligned
            if (threadIdx.x == 0) sm.chunk_n = (int)(hi - lo);
            w.res_mode = w.fresh = true;
            __syncthreads();
        }
    }
    if (w.res_mode && !w.in_smem && sm.chunk_n <= kChunkCap) {
        const int cn = sm.chunk_n;
        const uint16_t *src = w.g;
        for (int i = threadIdx.x * 8; i + 8 <= cn; i += kTPB * 8)
            *reinterpret_cast<uint4 *>(chunk_smem + i) = *reinterpret_cast<const uint4 *>(src + i);
        for (int i = (cn & ~7) + threadIdx.x; i < cn; i += kTPB) chunk_smem[i] = src[i];
        w.in_smem = true;
        __syncthreads();
    }
}

// End of the run: back to one contiguous stream in tok[step & 1] (step = merges done), where the host expects it.
template <class Grid>
__device__ __forceinline__ void resident_write_back(Grid &grid, const TrainView &v, uint32_t step, int cn, const uint16_t *chunk,
                                                    bool from_start) {
    if (threadIdx.x == 0) v.cta_counts[blockIdx.x] = (uint32_t)cn;
    grid.sync();
    unsigned long long pre = 0, all = 0;
    for (uint32_t j = threadIdx.x; j < gridDim.x; j += kTPB) {
        const unsigned long long c = __ldcg(&v.cta_counts[j]);
        all += c;
        if (j < blockIdx.x) pre += c;
    }
    __shared__ unsigned long long s_pre[kTPB / 32], s_all[kTPB / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        pre += __shfl_xor_sync(0xffffffffu, pre, o);
        all += __shfl_xor_sync(0xffffffffu, all, o);
    }
    if ((threadIdx.x & 31) == 0) { s_pre[threadIdx.x >> 5] = pre; s_all[threadIdx.x >> 5] = all; }
    __syncthreads();
    pre = all = 0;
    for (int q = 0; q < kTPB / 32; q++) { pre += s_pre[q]; all += s_all[q]; }
    // chunks that were formed before the first step may still live in buffer 0: gather in buffer 1 first
    const uint32_t to = from_start ? 1u : (step & 1u);
    uint16_t *dst = v.tok[to] + pre;
    for (int i = threadIdx.x; i < cn; i += kTPB) dst[i] = chunk[i];
    if (to != (step & 1u)) {
        grid.sync();
        const uint16_t *src = v.tok[1];
        uint16_t *fin = v.tok[0];
        const unsigned long long gt = (unsigned long long)blockIdx.x * kTPB + threadIdx.x, gn = (unsigned long long)gridDim.x * kTPB;
        for (unsigned long long i = gt * 8; i + 8 <= all; i += gn * 8)
            *reinterpret_cast<uint4 *>(fin + i) = *reinterpret_cast<const uint4 *>(src + i);
        for (unsigned long long i = (all & ~7ull) + gt; i < all; i += gn) fin[i] = src[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) v.dev->n[step & 1] = all;
}

__global__ void __launch_bounds__(kTPB, kCtasPerSm) train_loop_kernel(TrainView v, uint32_t n_steps) {
    cg::grid_group grid = cg::this_grid();
    __shared__ MergeSmem sm;
    __shared__ Best s_best;
    extern __shared__ __align__(16) uint16_t chunk[];  // kChunkCap tokens (resident tail)
    ChunkWhere w;
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    ECGB_MARK_DECL;
    uint32_t step = 0;
    for (; step < n_steps; step++) {
        ECGB_MARK_RESET;
        chunk_step_begin(v, step, sm, chunk, w);
        const Best fin = grid_argmax(grid, v, &s_best);
        if (gtid == 0) {
            v.best[step] = fin;
            if (fin.count == 0) atomicMin(&v.dev->done_step, step);
        }
        if (fin.count == 0) break;  // no pair left (lib.rs:88-90); uniform over the grid
        if (threadIdx.x == 0) atomicAdd(v.arrive, 1u);  // this CTA no longer reads the histogram in this step
        ECGB_MARK(0);
        if (w.res_mode) {
            const bool xx = (fin.key >> 16) == (fin.key & 0xFFFFu);
            if (xx || w.fresh) {  // otherwise the records published by the previous pass are all that is needed
                if (threadIdx.x < 32)
                    chunk_boundary(w.in_smem ? chunk : w.g, sm.chunk_n, fin.key >> 16, xx, &v.cta_bd[(step & 1) * v.cta_stride + blockIdx.x]);
                grid.sync();  // every chunk's boundary record is visible
            }
            w.fresh = false;
            const Boundary *bd = v.cta_bd + (step & 1) * v.cta_stride;
            if (w.in_smem) merge_pass<false, true>(v, step, fin, bd, v.main, sm, chunk);
#ifdef ECGB_TILE_RING
            else if (w.from_start) merge_pass<false, true, true>(v, step, fin, bd, v.main, sm, w.g, nullptr, &w.ring);
#endif
            else merge_pass<false, true>(v, step, fin, bd, v.main, sm, w.g);
        } else {
            merge_pass<false, false>(v, step, fin, nullptr, v.main, sm, nullptr);
        }
        ECGB_MARK_RESET;
        grid.sync();
        ECGB_MARK(10);
    }
    if (w.res_mode) resident_write_back(grid, v, step, sm.chunk_n, w.in_smem ? chunk : w.g, w.from_start);
}

// ------------------------------------------------------------------ persistent sharded loop

__device__ __forceinline__ uint32_t ld_vol(const uint32_t *p) { return *reinterpret_cast<const volatile uint32_t *>(p); }

// Record of this rank's shard as the stream that step s reads, without the (x,x) run fields.  One thread;
// everything it reads was written by other CTAs of this grid before they arrived at the counter the caller
// has just observed.
__device__ void shard_record(const TrainView &v, uint32_t s, bool resident, Boundary *out) {
    Boundary bd;
    memset(&bd, 0, sizeof(bd));
    for (int i = 0; i < 3; i++) bd.first[i] = kSentinel;
    bd.last[0] = bd.last[1] = kSentinel;
    if (!resident) {
        const volatile uint16_t *tok = v.tok[s & 1];
        const unsigned long long n = *reinterpret_cast<const volatile unsigned long long *>(&v.dev->n[s & 1]);
        bd.n_lo = (uint32_t)n;
        bd.n_hi = (uint32_t)(n >> 32);
        for (int i = 0; i < 3 && (unsigned long long)i < n; i++) bd.first[i] = tok[i];
        if (n >= 1) bd.last[1] = tok[n - 1];
        if (n >= 2) bd.last[0] = tok[n - 2];
    } else {
        const Boundary *cb = v.cta_bd + (size_t)(s & 1) * v.cta_stride;
        const unsigned long long n = *reinterpret_cast<const volatile unsigned long long *>(&v.n_hist[s]);  // sum of the chunk lengths
        bd.n_lo = (uint32_t)n;
        bd.n_hi = (uint32_t)(n >> 32);
        int nf = 0;
        for (int c = 0; c < (int)gridDim.x && nf < 3; c++) {
            const uint32_t nc = ld_vol(&cb[c].n_lo);
            for (int i = 0; i < 3 && (uint32_t)i < nc && nf < 3; i++) bd.first[nf++] = ld_vol(&cb[c].first[i]);
        }
        uint32_t got[2];
        int ng = 0;
        for (int c = (int)gridDim.x - 1; c >= 0 && ng < 2; c--) {
            const uint32_t nc = ld_vol(&cb[c].n_lo);
            if (nc >= 1) got[ng++] = ld_vol(&cb[c].last[1]);
            if (nc >= 2 && ng < 2) got[ng++] = ld_vol(&cb[c].last[0]);
        }
        if (ng >= 1) bd.last[1] = got[0];
        if (ng >= 2) bd.last[0] = got[1];
    }
    *out = bd;
}

// The (x,x) run fields of this rank's record for the pair (a, a): parity of the run of a that ends the shard
// and whether the shard is nothing else.  Warp 0 of one CTA; valid in lane 0.
__device__ void shard_run_fields(const TrainView &v, uint32_t s, bool resident, uint32_t a, uint32_t *trail_par, uint32_t *all_a) {
    const int lane = threadIdx.x & 31;
    if (resident) {  // fold the chunks' fields from the end of the shard
        if (lane == 0) {
            const Boundary *cb = v.cta_bd + (size_t)(s & 1) * v.cta_stride;
            uint32_t par = 0;
            bool all = true;
            for (int c = (int)gridDim.x - 1; c >= 0; c--) {
                const uint32_t nc = ld_vol(&cb[c].n_lo);
                if (nc == 0) continue;
                if (ld_vol(&cb[c].all_a)) { par ^= nc & 1u; continue; }
                par ^= ld_vol(&cb[c].trail_par);
                all = false;
                break;
            }
            *trail_par = par;
            *all_a = all ? 1u : 0u;
        }
        return;
    }
    const uint16_t *tok = v.tok[s & 1];
    const long long n = (long long)v.dev->n[s & 1];
    bool hit = false;
    long long found = -1;
    for (long long p = n - 1; p >= 0 && !hit; p -= 32) {
        const long long q = p - lane;
        const bool nonx = q >= 0 && tok[q] != a;
        const unsigned m = __ballot_sync(0xffffffffu, nonx);
        if (m) { found = p - (__ffs(m) - 1); hit = true; }
    }
    const long long run = n - 1 - found;
    *trail_par = (uint32_t)(run & 1);
    *all_a = run == n ? 1u : 0u;
}

// byte_pair_encoding (lib.rs:85-117) over a corpus cut into contiguous shards, one persistent cooperative
// kernel per rank (GPU), all running at the same time.  Per merge step a rank takes the argmax from ITS copy
// of the global pair histogram (the copies are identical, so is the winner), merges the pair in its shard
// -- the halo across shard borders comes from the peers' 64-byte shard records -- and writes its histogram
// patches both into its own table and, entry by entry over NVLink, into its inbox on every peer; the last CTA
// to finish publishes the shard record of the next stream and raises the step's flag on the peers.  Then
// every rank applies the patches the peers left in its area.  One exchange per step ((x,x) steps add one
// for the run parities), no launches, no host, no collective library inside the loop.
__device__ __forceinline__ void dist_loop_body(const TrainView &v, const PeerView &pv, uint32_t n_steps, MergeSmem &sm, Best &s_best,
                                               int &s_last, uint16_t *chunk) {
    AbortableGrid grid{v.gbar, v.abort, 0u};
    volatile unsigned int *abort = v.abort;
    ChunkWhere w;
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t gthreads = (uint64_t)gridDim.x * blockDim.x;
    uint8_t *const mine = pv.area[pv.rank];
    ECGB_MARK_DECL;
    uint32_t step = 0;
    for (; step < n_steps; step++) {
        ECGB_MARK_RESET;
        if (*abort) break;
        chunk_step_begin(v, step, sm, chunk, w);
        const Best fin = grid_argmax(grid, v, &s_best);
        if (gtid == 0) {
            v.best[step] = fin;
            if (fin.count == 0) atomicMin(&v.dev->done_step, step);
        }
        if (fin.count == 0) break;  // no pair left (lib.rs:88-90); the same on every rank
        if (threadIdx.x == 0) atomicAdd(v.arrive, 1u);
        const int par = (int)(step & 1u);
        const uint32_t a = fin.key >> 16;
        const bool xx = a == (fin.key & 0xFFFFu);
        if (w.res_mode && (xx || w.fresh)) {
            if (threadIdx.x < 32) chunk_boundary(w.in_smem ? chunk : w.g, sm.chunk_n, a, xx, &v.cta_bd[(size_t)par * v.cta_stride + blockIdx.x]);
            grid.sync();
        }
        w.fresh = false;
        if (xx && blockIdx.x == 0 && threadIdx.x < 32) {
            // second exchange of an (x,x) step: the run fields of this shard's record
            uint32_t tp = 0, alla = 0;
            shard_run_fields(v, step, w.res_mode, a, &tp, &alla);
            tp = __shfl_sync(0xffffffffu, tp, 0);
            alla = __shfl_sync(0xffffffffu, alla, 0);
            if ((int)threadIdx.x < pv.world && (int)threadIdx.x != pv.rank)  // lane r serves peer r
                st_unit(pv.area[threadIdx.x] + peer_rec_off(pv.world, par, pv.rank) + 8 * 5, tp | (alla << 1), peer_tag(pv.epoch, step));
        }
        ECGB_MARK(0);
        if (w.res_mode) {
            const Boundary *bd = v.cta_bd + (size_t)par * v.cta_stride;
            if (w.in_smem) merge_pass<false, true>(v, step, fin, bd, v.main, sm, chunk, &pv);
#ifdef ECGB_TILE_RING
            else if (w.from_start) merge_pass<false, true, true>(v, step, fin, bd, v.main, sm, w.g, &pv, &w.ring);
#endif
            else merge_pass<false, true>(v, step, fin, bd, v.main, sm, w.g, &pv);
        } else {
            merge_pass<false, false>(v, step, fin, nullptr, v.main, sm, nullptr, &pv);
        }
        ECGB_MARK_RESET;
        // this CTA's patches are in the local table (device-scope fence: the last CTA to arrive reads the list
        // length and the chunk records) and on their way to the peers (self-validating, no fence)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int old = atomicAdd(pv.done_ctas, 1u);
            s_last = old == (step + 1u) * gridDim.x - 1u;
        }
        __syncthreads();
        ECGB_MARK(11);
        if (s_last && threadIdx.x < 32) {
            // the whole rank has finished the pass: shard record of the next stream, list length, flag (warp 0)
            const int lane = threadIdx.x;
            __threadfence();
            Boundary *rec = &sm.bd_near[0];  // free between passes
            bool fast = false;
            if (w.res_mode) {
                // usual case: the first chunk holds >= 3 tokens and the last one >= 2 -- both records in one round trip
                const Boundary *cb = v.cta_bd + (size_t)(par ^ 1) * v.cta_stride;
                const uint32_t val = ld_vol(reinterpret_cast<const uint32_t *>(lane < 16 ? &cb[0] : &cb[gridDim.x - 1]) + (lane & 15));
                const unsigned long long n = *reinterpret_cast<const volatile unsigned long long *>(&v.n_hist[step + 1u]);
                const uint32_t n0 = __shfl_sync(0xffffffffu, val, 0), nl = __shfl_sync(0xffffffffu, val, 16);
                const uint32_t l0 = __shfl_sync(0xffffffffu, val, 16 + 5), l1 = __shfl_sync(0xffffffffu, val, 16 + 6);
                fast = n0 >= 3 && nl >= 2;
                if (fast && lane < 16) {
                    uint32_t w = 0;
                    if (lane == 0) w = (uint32_t)n;
                    else if (lane == 1) w = (uint32_t)(n >> 32);
                    else if (lane <= 4) w = val;  // first[0..2] of chunk 0
                    else if (lane == 5) w = l0;
                    else if (lane == 6) w = l1;
                    reinterpret_cast<uint32_t *>(rec)[lane] = w;
                }
            }
            if (!fast && lane == 0) shard_record(v, step + 1u, w.res_mode, rec);
            __syncwarp();
            uint32_t cnt = ld_vol(pv.out_count + par);
            if (cnt > pv.cap) cnt = pv.cap;  // overflow was flagged by push_entry
            if (lane == 0) pv.out_count[par ^ 1] = 0u;  // nobody appends before the barrier that ends this step
            if (lane < 5) {  // record of the stream step + 1 reads, unit `lane`, to every peer
                const uint32_t pay = lane == 0 ? rec->n_lo : lane == 1 ? rec->n_hi : lane == 2 ? (rec->first[0] | (rec->first[1] << 16))
                                   : lane == 3 ? (rec->first[2] | (rec->last[0] << 16)) : rec->last[1];
                const uint32_t tag_next = peer_tag(pv.epoch, step + 1u);
                for (int r = 0; r < pv.world; r++)
                    if (r != pv.rank) st_unit(pv.area[r] + peer_rec_off(pv.world, par ^ 1, pv.rank) + 8 * lane, pay, tag_next);
            }
            if (lane < pv.world && lane != pv.rank)  // lane r serves peer r: the step's flag with the list length
                st_unit(pv.area[lane] + peer_flag_off(pv.world, par, pv.rank), cnt, peer_tag(pv.epoch, step));
        }
        // the peers' patches of this step: kApplyCtas CTAs per peer wait for that peer's flag and apply its list, the
        // other CTAs go straight to the barrier (every CTA polling every peer's flag -- thousands of threads spinning on
        // a handful of words that NVLink writes have yet to land in -- was measurably slower)
        {
            const int n_peers = pv.world - 1;
            const int per_peer = max(1, min(kApplyCtas, (int)gridDim.x / max(n_peers, 1)));
            const int groups = max(1, (int)gridDim.x / per_peer);  // groups of per_peer CTAs; a group may serve several peers
            const int group = (int)blockIdx.x / per_peer, sub = (int)blockIdx.x % per_peer;
            for (int slot = group; group < groups && slot < n_peers; slot += groups) {
                const int r = slot < pv.rank ? slot : slot + 1;
                __syncthreads();  // peer_cnt[0] of the previous peer has been read
                if (threadIdx.x == 0)
                    sm.peer_cnt[0] = min(wait_unit(mine + peer_flag_off(pv.world, par, r), peer_tag(pv.epoch, step), abort, pv.timeout_ns,
                                                   (2u << 28) | ((uint32_t)r << 20) | (step & 0xFFFFFu)), pv.cap);
                __syncthreads();
                ECGB_MARK(12);
                const unsigned long long tau_val = sm.tau_val;  // read by the pass; tau only changes inside the argmax
                const uint32_t tag = peer_tag(pv.epoch, step);
                const uint32_t cnt = sm.peer_cnt[0];
                const uint4 *ent = reinterpret_cast<const uint4 *>(mine + peer_ent_off(pv.world, pv.cap, par, r));
                for (uint32_t i = (uint32_t)sub * kTPB + threadIdx.x; i < cnt; i += (uint32_t)per_peer * kTPB) {
                    // an entry may still be in flight behind the flag: both halves carry the tag
                    uint4 e;
                    unsigned long long t0 = 0;
                    for (unsigned int spins = 0;; spins++) {
                        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w) : "l"(ent + i) : "memory");
                        if (e.y == tag && e.w == tag) break;
                        if ((spins & 255u) == 255u) {
                            if (*abort) break;
                            unsigned long long now;
                            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                            if (t0 == 0) t0 = now;
                            else if (now - t0 > pv.timeout_ns) { *abort = 0x80000000u | (3u << 28) | ((uint32_t)r << 20) | (step & 0xFFFFFu); break; }
                        }
                    }
                    table_add(v.main, e.x, (long long)(int)e.z, tau_val);
                }
            }
        }
        ECGB_MARK(13);
        grid.sync();
        ECGB_MARK(10);
    }
    if (w.res_mode) resident_write_back(grid, v, step, sm.chunk_n, w.in_smem ? chunk : w.g, w.from_start);
    if (gtid == 0) v.dev->cur_step = step;
}

__global__ void __launch_bounds__(kTPB, kCtasPerSm) dist_loop_kernel(TrainView v, PeerView pv, uint32_t n_steps) {
    __shared__ MergeSmem sm;
    __shared__ Best s_best;
    __shared__ int s_last;
    extern __shared__ __align__(16) uint16_t chunk[];  // kChunkCap tokens (resident tail)
    dist_loop_body(v, pv, n_steps, sm, s_best, s_last, chunk);
}

// Every rank of a run in ONE cooperative launch on one device (blockIdx.y = rank): the co-residency the ranks'
// spin waits rely on is then guaranteed by the launch itself.  Same body, same protocol; used to exercise the
// exchange on a single GPU (tests).
struct DistLaunch {
    TrainView v;
    PeerView pv;
};
__global__ void __launch_bounds__(kTPB, kCtasPerSm) dist_loop_local_kernel(const DistLaunch *__restrict__ ranks, uint32_t n_steps) {
    __shared__ MergeSmem sm;
    __shared__ Best s_best;
    __shared__ int s_last;
    __shared__ DistLaunch s_me;
    extern __shared__ __align__(16) uint16_t chunk[];
    for (uint32_t i = threadIdx.x; i < sizeof(DistLaunch) / 4; i += kTPB)
        reinterpret_cast<uint32_t *>(&s_me)[i] = reinterpret_cast<const uint32_t *>(&ranks[blockIdx.y])[i];
    __syncthreads();
    dist_loop_body(s_me.v, s_me.pv, n_steps, sm, s_best, s_last, chunk);
}

// Sharded training: the argmax of one step as a cooperative launch (candidate list instead of a full
// table scan), then this shard's boundary record for the winning pair, as argmax_kernel leaves it.
__global__ void __launch_bounds__(kTPB, kCtasPerSm) dist_argmax_kernel(TrainView v, uint32_t step) {
    cg::grid_group grid = cg::this_grid();
    __shared__ Best s_best;
    if (step == kStepFromDevice) step = v.dev->cur_step;
    if (step > v.dev->max_steps) return;  // uniform over the grid
    const Best fin = grid_argmax(grid, v, &s_best);
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    // boundary record of this shard for the winning pair (warp 0 of CTA 0)
    const uint16_t *tok = v.tok[step & 1];
    const unsigned long long n = v.dev->n[step & 1];
    const uint32_t a = fin.key >> 16, bb = fin.key & 0xFFFFu;
    const int lane = threadIdx.x;
    unsigned long long run = 0;
    const bool xx = fin.count != 0 && a == bb && v.world > 1;
    if (xx) {  // trailing run of a, 32 tokens per probe
        bool hit = false;
        long long found = -1;
        for (long long p = (long long)n - 1; p >= 0 && !hit; p -= 32) {
            const long long q = p - lane;
            const bool nonx = q >= 0 && tok[q] != a;
            const unsigned m = __ballot_sync(0xffffffffu, nonx);
            if (m) { found = p - (__ffs(m) - 1); hit = true; }
        }
        run = (unsigned long long)((long long)n - 1 - found);
    }
    if (lane == 0) {
        v.best[step] = fin;
        if (fin.count == 0) atomicMin(&v.dev->done_step, step);
        Boundary bd;
        memset(&bd, 0, sizeof(bd));
        bd.n_lo = (uint32_t)n;
        bd.n_hi = (uint32_t)(n >> 32);
        for (int i = 0; i < 3; i++) bd.first[i] = (unsigned long long)i < n ? (uint32_t)tok[i] : kSentinel;
        bd.last[1] = n >= 1 ? (uint32_t)tok[n - 1] : kSentinel;
        bd.last[0] = n >= 2 ? (uint32_t)tok[n - 2] : kSentinel;
        if (xx) {
            bd.trail_par = (uint32_t)(run & 1);
            bd.all_a = run == n ? 1u : 0u;
        }
        *v.boundary = bd;
    }
}

// ------------------------------------------------------------------ delta lists (sharded training)

// list layout (u32 words): [0] count, [1] overflow, [2..3] pad, then cap keys, then cap 64-bit deltas
__global__ void compact_delta_kernel(PairTable d, uint32_t *list, uint32_t cap) {
    uint32_t *keys = list + 4;
    long long *deltas = reinterpret_cast<long long *>(list + 4 + cap);
    const uint64_t n = (uint64_t)d.mask + 1;
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t k = d.keys[s];
        if (k == kEmptyKey) continue;
        const long long c = (long long)d.cnt[s];
        d.keys[s] = kEmptyKey;
        d.cnt[s] = 0;
        if (c == 0) continue;
        const uint32_t at = atomicAdd(&list[0], 1u);
        if (at < cap) { keys[at] = k; deltas[at] = c; } else { list[1] = 1; }
    }
}

__global__ void advance_step_kernel(DevState *dev) { dev->cur_step++; }

__global__ void reset_list_kernel(uint32_t *list, uint32_t *used) { list[0] = 0; list[1] = 0; list[2] = 0; list[3] = 0; *used = 0; }

__global__ void apply_lists_kernel(PairTable main, const uint32_t *__restrict__ lists, uint32_t list_words, uint32_t cap,
                                   int world) {
    for (int r = 0; r < world; r++) {
        const uint32_t *list = lists + (size_t)r * list_words;
        const uint32_t cnt = min(list[0], cap);
        if (list[1]) atomicExch(main.overflow, 2u);
        const uint32_t *keys = list + 4;
        const long long *deltas = reinterpret_cast<const long long *>(list + 4 + cap);
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x)
            table_add(main, keys[i], deltas[i]);
    }
}

__global__ void widen_ids_kernel(const uint16_t *__restrict__ tok, uint32_t *__restrict__ out, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = tok[i];
}

}  // namespace ecgb

using namespace ecgb;

struct ecgb_trainer {
    int device = 0;
    int sms = 148;
    uint64_t capacity = 0;
    uint32_t max_merges = 0;
    uint32_t list_cap = 0;
    uint32_t steps_done = 0;   // merge steps applied so far
    uint32_t argmax_for = 0;   // best[] entries computed so far
    bool loaded = false;
    TrainView v{};
    void *blocks[32] = {nullptr};
    int n_blocks = 0;
    uint32_t *d_list = nullptr;  // this rank's delta list
    int coop_grid = 0;           // grid of dist_argmax_kernel (all CTAs co-resident)
    bool device_steps = false;   // steps were issued with ECGB_STEP_DEVICE
    // persistent sharded loop: this rank's receive area and its local counters
    uint8_t *peer_area = nullptr;
    uint64_t peer_bytes = 0;
    int peer_world = 0;
    uint32_t peer_cap = 0;
    uint32_t peer_epoch = 0;       // run number, part of every unit's tag
    uint32_t *peer_ctr = nullptr;  // out_count[2], done_ctas, overflow
};

static void print_phases() {
#ifdef ECGB_TRAIN_TIMING
    unsigned long long ph[16], zero[16] = {0};
    cudaMemcpyFromSymbol(ph, g_phase, sizeof(ph));
    cudaMemcpyToSymbol(g_phase, zero, sizeof(zero));
    static const char *names[14] = {"argmax", "pass setup", "tile load+sync", "flags+patches", "scan+sync", "stage",
                                    "look-back", "sync", "write-out", "patch flush", "grid sync", "fence+arrive",
                                    "wait peers", "apply"};
    for (int i = 0; i < 14; i++) fprintf(stderr, "phase %-14s %10.3f ms\n", names[i], ph[i] * 1e-6);
#endif
}

static int dev_alloc(ecgb_trainer *t, void **p, size_t bytes, bool zero) {
    if (t->n_blocks >= (int)(sizeof(t->blocks) / sizeof(t->blocks[0]))) return fail(ECGB_ECUDA, "trainer allocation table full");
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
    if (e != cudaSuccess) return fail(ECGB_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    t->blocks[t->n_blocks++] = *p;
    if (zero) {
        e = cudaMemset(*p, 0, bytes);
        if (e != cudaSuccess) return fail(ECGB_ECUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
    }
    return ECGB_OK;
}

static int alloc_table(ecgb_trainer *t, PairTable *pt, uint32_t log2cap) {
    const size_t cap = (size_t)1 << log2cap;
    int rc;
    if ((rc = dev_alloc(t, (void **)&pt->keys, cap * 4, false))) return rc;
    if ((rc = dev_alloc(t, (void **)&pt->cnt, cap * 8, true))) return rc;
    if ((rc = dev_alloc(t, (void **)&pt->used, 8, true))) return rc;
    pt->overflow = pt->used + 1;
    pt->mask = (uint32_t)(cap - 1);
    pt->tau = nullptr;
    pt->cand = pt->ncand = pt->inbits = nullptr;
    pt->push = nullptr;
    cudaError_t e = cudaMemset(pt->keys, 0xFF, cap * 4);
    if (e != cudaSuccess) return fail(ECGB_ECUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
    return ECGB_OK;
}

extern "C" int ecgb_trainer_create(int device, uint64_t capacity_tokens, uint32_t max_merges, uint32_t table_log2,
                                   ecgb_trainer **out) {
    ECGB_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    ECGB_REQUIRE(max_merges <= 0xFFFE - 256, "max_merges %u exceeds the 16-bit token id space", max_merges);
    ECGB_REQUIRE(capacity_tokens < (1ull << 41), "capacity too large");
    ECGB_REQUIRE(table_log2 == 0 || (table_log2 >= 10 && table_log2 <= 28), "table_log2 must be 0 (default) or in [10, 28]");
    int rc = check_device(device);
    if (rc) return rc;
    ecgb_trainer *t = new (std::nothrow) ecgb_trainer();
    if (!t) return fail(ECGB_ENOMEM, "host allocation failed");
    t->device = device;
    t->sms = sm_count(device);
    t->capacity = capacity_tokens;
    t->max_merges = max_merges;
    t->list_cap = 1u << 15;
    t->v.cta_stride = (uint32_t)(t->sms * kCtasPerSm);
    DeviceGuard g(device);
    const size_t tokbytes = (capacity_tokens + 64) * 2;
    const size_t ntiles = (size_t)(capacity_tokens / kTile) + 2;
    // every pair ever seen keeps its slot; measured on the 1.5e10-symbol corpus: 0.4 M / 1.0 M / 2.7 M slots after
    // 1 000 / 2 000 / 4 000 merges (about 28 M extrapolated at 20 000)
    if (table_log2 == 0) table_log2 = capacity_tokens >= (1ull << 31) ? 26 : capacity_tokens >= (1ull << 28) ? 24 : 22;
    rc = dev_alloc(t, (void **)&t->v.tok[0], tokbytes, false);
    if (!rc) rc = dev_alloc(t, (void **)&t->v.tok[1], tokbytes, false);
    if (!rc) rc = dev_alloc(t, (void **)&t->v.dev, sizeof(DevState), true);
    if (!rc) rc = alloc_table(t, &t->v.main, table_log2);
    if (!rc) rc = alloc_table(t, &t->v.delta, 18);
    if (!rc) {  // argmax candidate list of the main table
        const size_t cap = (size_t)1 << table_log2;
        unsigned long long *tau = nullptr;
        rc = dev_alloc(t, (void **)&t->v.main.cand, cap * 4, false);
        if (!rc) rc = dev_alloc(t, (void **)&t->v.main.inbits, cap / 8 + 8, true);
        if (!rc) rc = dev_alloc(t, (void **)&tau, 16, true);
        t->v.main.tau = tau;
        t->v.main.ncand = reinterpret_cast<uint32_t *>(tau + 1);
    }
    if (!rc) rc = dev_alloc(t, (void **)&t->v.best, sizeof(Best) * ((size_t)max_merges + 1), true);
    if (!rc) rc = dev_alloc(t, (void **)&t->v.partial, sizeof(Best) * ((size_t)t->v.cta_stride + 1), true);
    if (!rc) rc = dev_alloc(t, (void **)&t->v.tickets, 4 * ((size_t)max_merges + 1), true);
    if (!rc) rc = dev_alloc(t, (void **)&t->v.tile_status, 8 * ntiles, true);
    if (!rc) rc = dev_alloc(t, (void **)&t->v.boundary, sizeof(Boundary), true);
    if (!rc) rc = dev_alloc(t, (void **)&t->v.n_hist, 8 * ((size_t)max_merges + 2), true);
    if (!rc) rc = dev_alloc(t, (void **)&t->v.arrive, 16, true);
    t->v.gbar = t->v.arrive + 1;
    t->v.abort = t->v.arrive + 2;
    if (!rc) rc = dev_alloc(t, (void **)&t->v.cta_bd, sizeof(Boundary) * 2 * (size_t)t->v.cta_stride, true);
    if (!rc) rc = dev_alloc(t, (void **)&t->v.cta_counts, 4 * (size_t)t->v.cta_stride, true);
    if (!rc) rc = dev_alloc(t, (void **)&t->d_list, (size_t)(4 + 3 * (size_t)t->list_cap) * 4, true);
    if (rc) { ecgb_trainer_destroy(t); return rc; }
    t->v.rank = 0;
    t->v.world = 1;
    *out = t;
    return ECGB_OK;
}

extern "C" int ecgb_trainer_destroy(ecgb_trainer *t) {
    if (!t) return ECGB_OK;
    {
        DeviceGuard g(t->device);
        for (int i = 0; i < t->n_blocks; i++) cudaFree(t->blocks[i]);
    }
    delete t;
    return ECGB_OK;
}

static int reset_state(ecgb_trainer *t, uint64_t n, cudaStream_t st) {
    DevState h{};
    h.n[0] = n; h.n[1] = 0; h.done_step = 0xFFFFFFFFu; h.argmax_done = 0; h.cur_step = 0; h.max_steps = t->max_merges;
    ECGB_CUDA(cudaMemcpyAsync(t->v.dev, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    ECGB_CUDA(cudaStreamSynchronize(st));  // h is a stack object
    const size_t cap = (size_t)t->v.main.mask + 1;
    ECGB_CUDA(cudaMemsetAsync(t->v.main.keys, 0xFF, cap * 4, st));
    ECGB_CUDA(cudaMemsetAsync(t->v.main.cnt, 0, cap * 8, st));
    ECGB_CUDA(cudaMemsetAsync(t->v.main.used, 0, 8, st));
    // tau = 2^64 - 1: the first argmax finds an empty candidate list and builds it
    ECGB_CUDA(cudaMemsetAsync(const_cast<unsigned long long *>(t->v.main.tau), 0xFF, 8, st));
    ECGB_CUDA(cudaMemsetAsync(t->v.main.ncand, 0, 8, st));
    ECGB_CUDA(cudaMemsetAsync(t->v.main.inbits, 0, cap / 8 + 8, st));
    const size_t dcap = (size_t)t->v.delta.mask + 1;
    ECGB_CUDA(cudaMemsetAsync(t->v.delta.keys, 0xFF, dcap * 4, st));
    ECGB_CUDA(cudaMemsetAsync(t->v.delta.cnt, 0, dcap * 8, st));
    ECGB_CUDA(cudaMemsetAsync(t->v.delta.used, 0, 8, st));
    ECGB_CUDA(cudaMemsetAsync(t->v.tickets, 0, 4 * ((size_t)t->max_merges + 1), st));
    ECGB_CUDA(cudaMemsetAsync(t->v.n_hist, 0, 8 * ((size_t)t->max_merges + 2), st));
    ECGB_CUDA(cudaMemsetAsync(t->v.arrive, 0, 16, st));
    ECGB_CUDA(cudaMemcpyAsync(t->v.n_hist, &t->v.dev->n[0], 8, cudaMemcpyDeviceToDevice, st));
    ECGB_CUDA(cudaMemsetAsync(t->v.best, 0, sizeof(Best) * ((size_t)t->max_merges + 1), st));
    ECGB_CUDA(cudaMemsetAsync(t->v.tile_status, 0, 8 * ((size_t)(t->capacity / kTile) + 2), st));
    t->steps_done = 0;
    t->argmax_for = 0;
    t->device_steps = false;
    t->loaded = true;
    return ECGB_OK;
}

extern "C" int ecgb_trainer_load_device(ecgb_trainer *t, const uint8_t *d_text, uint64_t n, void *stream) {
    ECGB_REQUIRE(t, "trainer is NULL");
    ECGB_REQUIRE(n <= t->capacity, "corpus of %llu bytes exceeds the trainer capacity %llu", (unsigned long long)n,
                 (unsigned long long)t->capacity);
    ECGB_REQUIRE(n == 0 || d_text, "d_text is NULL");
    DeviceGuard g(t->device);
    cudaStream_t st = as_stream(stream);
    int rc = reset_state(t, n, st);
    if (rc) return rc;
    if (n) {
        bytes_to_tokens_kernel<<<t->sms * 8, 256, 0, st>>>(d_text, t->v.tok[0], n);
        ECGB_CUDA(cudaGetLastError());
    }
    return ECGB_OK;
}

extern "C" int ecgb_trainer_load_host(ecgb_trainer *t, const uint8_t *h_text, uint64_t n) {
    ECGB_REQUIRE(t, "trainer is NULL");
    ECGB_REQUIRE(n <= t->capacity, "corpus of %llu bytes exceeds the trainer capacity %llu", (unsigned long long)n,
                 (unsigned long long)t->capacity);
    ECGB_REQUIRE(n == 0 || h_text, "h_text is NULL");
    DeviceGuard g(t->device);
    // stage the bytes in the (not yet used) second token buffer
    uint8_t *d_bytes = reinterpret_cast<uint8_t *>(t->v.tok[1]);
    if (n) ECGB_CUDA(cudaMemcpy(d_bytes, h_text, n, cudaMemcpyHostToDevice));
    int rc = ecgb_trainer_load_device(t, d_bytes, n, nullptr);
    if (rc) return rc;
    ECGB_CUDA(cudaStreamSynchronize(0));
    return ECGB_OK;
}

static int check_tables(ecgb_trainer *t) {
    uint32_t flags[2] = {0, 0};
    ECGB_CUDA(cudaMemcpy(flags, t->v.main.used, 8, cudaMemcpyDeviceToHost));
    if (flags[1]) return fail(ECGB_ECAPACITY, "pair table overflow (code %u): %u of %u slots used; raise table_log2", flags[1],
                              flags[0], t->v.main.mask + 1);
    uint32_t dflags[2] = {0, 0};
    ECGB_CUDA(cudaMemcpy(dflags, t->v.delta.used, 8, cudaMemcpyDeviceToHost));
    if (dflags[1]) return fail(ECGB_ECAPACITY, "delta table overflow");
    uint32_t sync_words[3] = {0, 0, 0};  // arrive, gbar, abort
    ECGB_CUDA(cudaMemcpy(sync_words, t->v.arrive, 12, cudaMemcpyDeviceToHost));
    if (sync_words[2]) {
        const uint32_t w = sync_words[2];
        static const char *what[4] = {"?", "shard record", "step flag", "patch entry"};
        return fail(ECGB_ECUDA, "sharded training: the wait for a %s (unit %u) of rank %u in step %u timed out -- is every rank running?",
                    what[(w >> 28) & 3u], (w >> 24) & 15u, (w >> 20) & 15u, w & 0xFFFFFu);
    }
    if (t->peer_ctr) {
        uint32_t ctr[4] = {0, 0, 0, 0};
        ECGB_CUDA(cudaMemcpy(ctr, t->peer_ctr, 16, cudaMemcpyDeviceToHost));
        if (ctr[3]) return fail(ECGB_ECAPACITY, "sharded training: a step produced more than %u histogram patches per rank", t->peer_cap);
    }
    return ECGB_OK;
}

extern "C" int ecgb_trainer_run(ecgb_trainer *t, uint32_t num_merges, uint32_t *h_pairs, uint64_t *h_counts,
                                uint32_t *h_ntied, uint32_t *n_done) {
    ECGB_REQUIRE(t && n_done, "NULL argument");
    *n_done = 0;
    ECGB_REQUIRE(t->loaded, "no corpus loaded");
    ECGB_REQUIRE(t->v.world == 1, "ecgb_trainer_run is the single-device loop; use the dist_* calls for shards");
    ECGB_REQUIRE(t->steps_done == 0, "trainer already ran; load the corpus again");
    ECGB_REQUIRE(num_merges <= t->max_merges, "num_merges %u > max_merges %u", num_merges, t->max_merges);
    DeviceGuard g(t->device);
    cudaStream_t st = 0;
    uint64_t n0 = 0;
    ECGB_CUDA(cudaMemcpy(&n0, &t->v.dev->n[0], 8, cudaMemcpyDeviceToHost));
    // get_stats once; afterwards the histogram is patched by the merge passes
    count_kernel<<<t->sms * 4, 256, 0, st>>>(t->v.tok[0], n0, kSentinel, t->v.main);
    ECGB_CUDA(cudaGetLastError());
    // one persistent cooperative kernel runs every step (grid = all co-resident CTAs)
    int per_sm = 0;
    const size_t dyn_smem = (size_t)kChunkCap * sizeof(uint16_t);
    ECGB_CUDA(cudaFuncSetAttribute(train_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    ECGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, train_loop_kernel, kTPB, dyn_smem));
    if (per_sm < 1) return fail(ECGB_EUNSUPPORTED, "train_loop_kernel does not fit on this device");
    int grid = t->sms * std::min(per_sm, kCtasPerSm);
    if (grid > (int)t->v.cta_stride) grid = (int)t->v.cta_stride;
    if (num_merges > 0) {
        TrainView view = t->v;
        const char *knob = getenv("ECGB_REDUNDANT_ARGMAX");  // tuning knob (profiles/train_knobs.py)
        view.redundant_max = knob ? (uint32_t)atoi(knob) : kRedundantArgmax;
        knob = getenv("ECGB_RESIDENT_TAIL");
        view.resident_ok = knob ? (uint32_t)atoi(knob) : 2u;
        uint32_t steps = num_merges;
        void *kargs[] = {&view, &steps};
        ECGB_CUDA(cudaLaunchCooperativeKernel((const void *)train_loop_kernel, dim3(grid), dim3(kTPB), kargs, dyn_smem, st));
    }
    ECGB_CUDA(cudaGetLastError());
    ECGB_CUDA(cudaStreamSynchronize(st));
    print_phases();
    int rc = check_tables(t);
    if (rc) return rc;
    DevState hs;
    ECGB_CUDA(cudaMemcpy(&hs, t->v.dev, sizeof(hs), cudaMemcpyDeviceToHost));
    const uint32_t done = std::min(num_merges, hs.done_step);
    t->steps_done = done;
    t->argmax_for = num_merges;
    std::vector<Best> hb(done ? done : 1);
    if (done) ECGB_CUDA(cudaMemcpy(hb.data(), t->v.best, sizeof(Best) * done, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < done; i++) {
        if (h_pairs) { h_pairs[2 * i] = hb[i].key >> 16; h_pairs[2 * i + 1] = hb[i].key & 0xFFFFu; }
        if (h_counts) h_counts[i] = hb[i].count;
        if (h_ntied) h_ntied[i] = hb[i].ntied;
    }
    *n_done = done;
    return ECGB_OK;
}

extern "C" int ecgb_trainer_length(ecgb_trainer *t, uint64_t *n_out) {
    ECGB_REQUIRE(t && n_out, "NULL argument");
    DeviceGuard g(t->device);
    ECGB_CUDA(cudaDeviceSynchronize());
    DevState hs;
    ECGB_CUDA(cudaMemcpy(&hs, t->v.dev, sizeof(hs), cudaMemcpyDeviceToHost));
    *n_out = hs.n[t->steps_done & 1];
    return ECGB_OK;
}

extern "C" int ecgb_trainer_ids_host(ecgb_trainer *t, uint32_t *h_ids, uint64_t cap, uint64_t *n_out) {
    ECGB_REQUIRE(t && n_out, "NULL argument");
    uint64_t n = 0;
    int rc = ecgb_trainer_length(t, &n);
    if (rc) return rc;
    *n_out = n;
    if (n > cap) return fail(ECGB_ECAPACITY, "ids buffer too small: need %llu", (unsigned long long)n);
    if (n == 0) return ECGB_OK;
    ECGB_REQUIRE(h_ids, "h_ids is NULL");
    DeviceGuard g(t->device);
    // widen in chunks through the idle token buffer
    uint32_t *d_tmp = reinterpret_cast<uint32_t *>(t->v.tok[(t->steps_done + 1) & 1]);
    const uint64_t chunk_cap = (t->capacity + 64) / 2;
    const uint16_t *src = t->v.tok[t->steps_done & 1];
    for (uint64_t o = 0; o < n; o += chunk_cap) {
        const uint64_t c = std::min(chunk_cap, n - o);
        widen_ids_kernel<<<t->sms * 8, 256>>>(src + o, d_tmp, c);
        ECGB_CUDA(cudaGetLastError());
        ECGB_CUDA(cudaMemcpy(h_ids + o, d_tmp, c * 4, cudaMemcpyDeviceToHost));
    }
    return ECGB_OK;
}

// Length of this shard's token stream before merge step i, i in [0, n_steps]
// (h_n[0] = corpus bytes, h_n[i] = tokens left after i merges): the n_t of the
// algorithmic-bytes formula 2*(n_t + n_{t+1}) per step.
extern "C" int ecgb_trainer_lengths(ecgb_trainer *t, uint32_t n_steps, uint64_t *h_n) {
    ECGB_REQUIRE(t && h_n, "NULL argument");
    ECGB_REQUIRE(n_steps <= t->max_merges, "n_steps out of range");
    DeviceGuard g(t->device);
    ECGB_CUDA(cudaDeviceSynchronize());
    ECGB_CUDA(cudaMemcpy(h_n, t->v.n_hist, 8 * ((size_t)n_steps + 1), cudaMemcpyDeviceToHost));
    return ECGB_OK;
}

// Apply a given list of merges, in order, to the loaded text: merge (lib.rs:10-26) once per pair, the id each
// pair becomes given by the caller -- what the reference's track_encoding (tokenizer_utils.py:95-134) does with
// pair-form merges.  No histogram, no argmax.  Read the result with ecgb_trainer_ids_host.
extern "C" int ecgb_trainer_apply_pairs(ecgb_trainer *t, const uint32_t *h_pairs, const uint32_t *h_new_ids, uint32_t n) {
    ECGB_REQUIRE(t, "trainer is NULL");
    ECGB_REQUIRE(t->loaded && t->steps_done == 0 && !t->device_steps, "load the text first");
    ECGB_REQUIRE(n <= t->max_merges, "%u merges > max_merges %u", n, t->max_merges);
    ECGB_REQUIRE(n == 0 || (h_pairs && h_new_ids), "NULL argument");
    for (uint32_t i = 0; i < n; i++)
        ECGB_REQUIRE(h_pairs[2 * i] < kSentinel && h_pairs[2 * i + 1] < kSentinel && h_new_ids[i] < kSentinel,
                     "merge %u: ids must be below 65535", i);
    if (n == 0) return ECGB_OK;
    DeviceGuard g(t->device);
    cudaStream_t st = 0;
    std::vector<Best> hb(n);
    for (uint32_t i = 0; i < n; i++) hb[i] = Best{1ull, (h_pairs[2 * i] << 16) | h_pairs[2 * i + 1], 1u};
    ECGB_CUDA(cudaMemcpy(t->v.best, hb.data(), sizeof(Best) * n, cudaMemcpyHostToDevice));
    uint32_t *d_ids = nullptr;
    int rc = dev_alloc(t, (void **)&d_ids, 4 * (size_t)n, false);
    if (rc) return rc;
    ECGB_CUDA(cudaMemcpy(d_ids, h_new_ids, 4 * (size_t)n, cudaMemcpyHostToDevice));
    TrainView view = t->v;
    view.new_ids = d_ids;
    const int merge_grid = (int)std::min<uint64_t>((uint64_t)t->sms * 6, t->capacity / kTile + 1);
    for (uint32_t step = 0; step < n; step++) {
        merge_kernel<<<merge_grid, kTPB, 0, st>>>(view, step, nullptr, t->v.delta);
        if ((step & 255u) == 255u) {  // the patches are not wanted: keep the delta table from filling up
            const size_t dcap = (size_t)t->v.delta.mask + 1;
            ECGB_CUDA(cudaMemsetAsync(t->v.delta.keys, 0xFF, dcap * 4, st));
            ECGB_CUDA(cudaMemsetAsync(t->v.delta.cnt, 0, dcap * 8, st));
            ECGB_CUDA(cudaMemsetAsync(t->v.delta.used, 0, 8, st));
        }
    }
    ECGB_CUDA(cudaGetLastError());
    ECGB_CUDA(cudaStreamSynchronize(st));
    const size_t dcap = (size_t)t->v.delta.mask + 1;
    ECGB_CUDA(cudaMemset(t->v.delta.keys, 0xFF, dcap * 4));
    ECGB_CUDA(cudaMemset(t->v.delta.cnt, 0, dcap * 8));
    ECGB_CUDA(cudaMemset(t->v.delta.used, 0, 8));
    t->steps_done = n;
    return ECGB_OK;
}

// Occupancy of the pair table: h_out[0] = slots claimed (keys are never released: a pair whose count fell to
// zero keeps its slot), [1] = capacity, [2] = argmax candidates listed, [3] = overflow flag.
extern "C" int ecgb_trainer_table_stats(ecgb_trainer *t, uint64_t h_out[4]) {
    ECGB_REQUIRE(t && h_out, "NULL argument");
    DeviceGuard g(t->device);
    ECGB_CUDA(cudaDeviceSynchronize());
    uint32_t flags[2] = {0, 0}, nc = 0;
    ECGB_CUDA(cudaMemcpy(flags, t->v.main.used, 8, cudaMemcpyDeviceToHost));
    ECGB_CUDA(cudaMemcpy(&nc, t->v.main.ncand, 4, cudaMemcpyDeviceToHost));
    h_out[0] = flags[0];
    h_out[1] = (uint64_t)t->v.main.mask + 1;
    h_out[2] = nc;
    h_out[3] = flags[1];
    return ECGB_OK;
}

// The live pair histogram (every pair with a non-zero count): get_stats (lib.rs:28-48)
// of the current token stream.  Used by the parity tests.
extern "C" int ecgb_trainer_histogram(ecgb_trainer *t, uint32_t *h_pairs, int64_t *h_counts, uint64_t cap,
                                      uint64_t *n_out) {
    ECGB_REQUIRE(t && n_out, "NULL argument");
    DeviceGuard g(t->device);
    ECGB_CUDA(cudaDeviceSynchronize());
    const size_t slots = (size_t)t->v.main.mask + 1;
    std::vector<uint32_t> keys(slots);
    std::vector<unsigned long long> cnt(slots);
    ECGB_CUDA(cudaMemcpy(keys.data(), t->v.main.keys, slots * 4, cudaMemcpyDeviceToHost));
    ECGB_CUDA(cudaMemcpy(cnt.data(), t->v.main.cnt, slots * 8, cudaMemcpyDeviceToHost));
    uint64_t k = 0;
    for (size_t i = 0; i < slots; i++) {
        if (keys[i] == kEmptyKey || cnt[i] == 0) continue;
        if (k < cap && h_pairs && h_counts) {
            h_pairs[2 * k] = keys[i] >> 16;
            h_pairs[2 * k + 1] = keys[i] & 0xFFFFu;
            h_counts[k] = (int64_t)cnt[i];
        }
        k++;
    }
    *n_out = k;
    if (k > cap) return fail(ECGB_ECAPACITY, "histogram has %llu entries", (unsigned long long)k);
    return ECGB_OK;
}

// ------------------------------------------------------------------ sharded (step-wise) interface

extern "C" int ecgb_trainer_dist_sizes(const ecgb_trainer *t, uint32_t *boundary_bytes, uint32_t *list_bytes) {
    ECGB_REQUIRE(t && boundary_bytes && list_bytes, "NULL argument");
    *boundary_bytes = kBoundaryWords * 4;
    *list_bytes = (4 + 3 * t->list_cap) * 4;
    return ECGB_OK;
}

// Declare this trainer to be shard `rank` of `world`, and publish its boundary record
// (no pair selected yet) to d_boundary_out.
extern "C" int ecgb_trainer_dist_begin(ecgb_trainer *t, int rank, int world, void *d_boundary_out, void *stream) {
    ECGB_REQUIRE(t && d_boundary_out, "NULL argument");
    ECGB_REQUIRE(t->loaded && t->steps_done == 0, "load the shard first");
    ECGB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
    t->v.rank = rank;
    t->v.world = world;
    DeviceGuard g(t->device);
    cudaStream_t st = as_stream(stream);
    // an argmax over the (still empty) table just writes the boundary record: count == 0 pair
    TrainView v = t->v;
    v.best = t->v.best + t->max_merges;  // scratch slot
    argmax_kernel<<<1, 256, 0, st>>>(v, 0);
    ECGB_CUDA(cudaGetLastError());
    // the scratch argmax recorded "done" because the table is empty: clear that
    const uint32_t none = 0xFFFFFFFFu;
    ECGB_CUDA(cudaMemcpyAsync(&t->v.dev->done_step, &none, 4, cudaMemcpyHostToDevice, st));
    ECGB_CUDA(cudaMemcpyAsync(d_boundary_out, t->v.boundary, sizeof(Boundary), cudaMemcpyDeviceToDevice, st));
    ECGB_CUDA(cudaStreamSynchronize(st));
    return ECGB_OK;
}

// Local get_stats of this shard (including the window that straddles into the next
// shard) as a delta list in d_list_out.
extern "C" int ecgb_trainer_dist_count(ecgb_trainer *t, const void *d_all_boundaries, void *d_list_out, void *stream) {
    ECGB_REQUIRE(t && d_all_boundaries && d_list_out, "NULL argument");
    DeviceGuard g(t->device);
    cudaStream_t st = as_stream(stream);
    // the right neighbour token is needed on the host side of the launch: read the gathered records
    std::vector<Boundary> all((size_t)t->v.world);
    ECGB_CUDA(cudaMemcpyAsync(all.data(), d_all_boundaries, sizeof(Boundary) * all.size(), cudaMemcpyDeviceToHost, st));
    ECGB_CUDA(cudaStreamSynchronize(st));
    uint32_t right = kSentinel;
    for (int r = t->v.rank + 1; r < t->v.world; r++) {
        const uint64_t n = ((uint64_t)all[r].n_hi << 32) | all[r].n_lo;
        if (n) { right = all[r].first[0]; break; }
    }
    const uint64_t n = ((uint64_t)all[t->v.rank].n_hi << 32) | all[t->v.rank].n_lo;
    uint32_t *list = static_cast<uint32_t *>(d_list_out);
    reset_list_kernel<<<1, 1, 0, st>>>(list, t->v.delta.used);
    count_kernel<<<t->sms * 4, 256, 0, st>>>(t->v.tok[0], n, right, t->v.delta);
    compact_delta_kernel<<<t->sms * 2, 256, 0, st>>>(t->v.delta, list, t->list_cap);
    ECGB_CUDA(cudaGetLastError());
    return ECGB_OK;
}

// Apply every rank's delta list to this rank's copy of the global histogram, take the
// argmax for `step` and publish this shard's boundary record for the winning pair.
extern "C" int ecgb_trainer_dist_commit(ecgb_trainer *t, uint32_t step, const void *d_all_lists, void *d_boundary_out,
                                        void *stream) {
    ECGB_REQUIRE(t && d_all_lists && d_boundary_out, "NULL argument");
    ECGB_REQUIRE(step == kStepFromDevice || step <= t->max_merges, "step %u out of range", step);
    ECGB_REQUIRE(step != kStepFromDevice || getenv("ECGB_DIST_FULLSCAN") == nullptr, "the full-scan argmax has no device-side step");
    DeviceGuard g(t->device);
    cudaStream_t st = as_stream(stream);
    const uint32_t list_words = 4 + 3 * t->list_cap;
    apply_lists_kernel<<<t->sms, 256, 0, st>>>(t->v.main, static_cast<const uint32_t *>(d_all_lists), list_words,
                                               t->list_cap, t->v.world);
    static const bool full_scan = getenv("ECGB_DIST_FULLSCAN") != nullptr;  // A/B knob: the old full-table argmax
    if (full_scan) {
        argmax_kernel<<<(int)t->v.cta_stride, 256, 0, st>>>(t->v, step);
    } else {
    if (t->coop_grid == 0) {
        int per_sm = 0;
        ECGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dist_argmax_kernel, kTPB, 0));
        if (per_sm < 1) return fail(ECGB_EUNSUPPORTED, "dist_argmax_kernel does not fit on this device");
        t->coop_grid = std::min(t->sms * std::min(per_sm, kCtasPerSm), (int)t->v.cta_stride);
    }
    {
        TrainView view = t->v;
        view.redundant_max = kRedundantArgmax;
        void *kargs[] = {&view, &step};
        ECGB_CUDA(cudaLaunchCooperativeKernel((const void *)dist_argmax_kernel, dim3(t->coop_grid), dim3(kTPB), kargs, 0, st));
    }
    }
    ECGB_CUDA(cudaGetLastError());
    ECGB_CUDA(cudaMemcpyAsync(d_boundary_out, t->v.boundary, sizeof(Boundary), cudaMemcpyDeviceToDevice, st));
    if (step == kStepFromDevice) t->device_steps = true;
    else t->argmax_for = step + 1;
    return ECGB_OK;
}

// Merge best[step] in this shard given every rank's boundary record; the histogram
// patches go to d_list_out.
extern "C" int ecgb_trainer_dist_merge(ecgb_trainer *t, uint32_t step, const void *d_all_boundaries, void *d_list_out,
                                       void *stream) {
    ECGB_REQUIRE(t && d_all_boundaries && d_list_out, "NULL argument");
    ECGB_REQUIRE(step == kStepFromDevice || step < t->max_merges, "step %u out of range", step);
    DeviceGuard g(t->device);
    cudaStream_t st = as_stream(stream);
    uint32_t *list = static_cast<uint32_t *>(d_list_out);
    reset_list_kernel<<<1, 1, 0, st>>>(list, t->v.delta.used);
    const int merge_grid = (int)std::min<uint64_t>((uint64_t)t->sms * 6, t->capacity / kTile + 1);
    merge_kernel<<<merge_grid, kTPB, 0, st>>>(t->v, step, static_cast<const Boundary *>(d_all_boundaries), t->v.delta);
    compact_delta_kernel<<<t->sms * 2, 256, 0, st>>>(t->v.delta, list, t->list_cap);
    ECGB_CUDA(cudaGetLastError());
    if (step == kStepFromDevice) t->device_steps = true;
    else t->steps_done = step + 1;
    return ECGB_OK;
}

// ------------------------------------------------------------------ persistent sharded loop (device-initiated exchange)

// Allocate (once) and clear this rank's receive area for a run over `world` ranks.  The peers need its address:
// inside one process the pointer itself, between processes an IPC handle (ecgb_ipc_export / ecgb_ipc_open).
// Clearing must be complete on EVERY rank before any rank calls ecgb_trainer_dist_run (host-side barrier).
extern "C" int ecgb_trainer_peer_area(ecgb_trainer *t, int world, void **d_area, uint64_t *bytes) {
    ECGB_REQUIRE(t && d_area && bytes, "NULL argument");
    ECGB_REQUIRE(world >= 1 && world <= kMaxWorld, "world %d out of range [1, %d]", world, kMaxWorld);
    DeviceGuard g(t->device);
    const uint32_t cap = 1u << 20;  // patch entries per (parity, source): 16 MB each
    const uint64_t need = peer_area_bytes(world, cap);
    if (t->peer_area == nullptr || t->peer_world != world) {
        ECGB_REQUIRE(t->peer_area == nullptr, "the receive area was created for %d ranks", t->peer_world);
        int rc = dev_alloc(t, (void **)&t->peer_area, need, true);
        if (!rc) rc = dev_alloc(t, (void **)&t->peer_ctr, 16, true);
        if (rc) return rc;
        t->peer_bytes = need;
        t->peer_world = world;
        t->peer_cap = cap;
    }
    // flags and records; list entries validate themselves through the run's epoch
    t->peer_epoch = (t->peer_epoch % 4095u) + 1u;
    ECGB_CUDA(cudaMemset(t->peer_area, 0, peer_ent_off(world, cap, 0, 0)));
    ECGB_CUDA(cudaMemset(t->peer_ctr, 0, 16));
    ECGB_CUDA(cudaDeviceSynchronize());
    *d_area = t->peer_area;
    *bytes = need;
    return ECGB_OK;
}

extern "C" int ecgb_ipc_export(const void *d_ptr, uint8_t handle[64]) {
    ECGB_REQUIRE(d_ptr && handle, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    ECGB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
    std::memcpy(handle, &h, 64);
    return ECGB_OK;
}

extern "C" int ecgb_ipc_open(const uint8_t handle[64], int device, void **d_ptr) {
    ECGB_REQUIRE(d_ptr && handle, "NULL argument");
    int rc = check_device(device);
    if (rc) return rc;
    DeviceGuard g(device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    ECGB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return ECGB_OK;
}

extern "C" int ecgb_ipc_close(void *d_ptr, int device) {
    if (!d_ptr) return ECGB_OK;
    DeviceGuard g(device);
    ECGB_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return ECGB_OK;
}

// Apply every rank's delta list to this rank's copy of the global histogram (the get_stats exchange that
// precedes ecgb_trainer_dist_run; ecgb_trainer_dist_commit without the argmax).
extern "C" int ecgb_trainer_dist_apply(ecgb_trainer *t, const void *d_all_lists, int world, void *stream) {
    ECGB_REQUIRE(t && d_all_lists, "NULL argument");
    ECGB_REQUIRE(world >= 1 && world <= kMaxWorld, "world %d out of range", world);
    DeviceGuard g(t->device);
    const uint32_t list_words = 4 + 3 * t->list_cap;
    apply_lists_kernel<<<t->sms, 256, 0, as_stream(stream)>>>(t->v.main, static_cast<const uint32_t *>(d_all_lists), list_words,
                                                              t->list_cap, world);
    ECGB_CUDA(cudaGetLastError());
    return ECGB_OK;
}

// Launch the persistent sharded loop of this rank (asynchronous; every rank of the run must launch, each on
// its own device or -- with max_ctas small enough for all of them to be co-resident -- on one device).
// d_areas[r] = rank r's receive area as addressable from this process; d_all_boundaries = the gathered
// records of ecgb_trainer_dist_begin.  The table must hold the global histogram (dist_count + dist_apply).
// Read the outcome with ecgb_trainer_results.
static int dist_prepare(ecgb_trainer *t, int rank, int world, void *const *d_areas, const void *d_all_boundaries,
                        uint32_t num_merges, double timeout_s, cudaStream_t st, TrainView *view_out, PeerView *pv_out) {
    ECGB_REQUIRE(t && d_areas && d_all_boundaries, "NULL argument");
    ECGB_REQUIRE(t->loaded && t->steps_done == 0 && !t->device_steps, "load the shard first");
    ECGB_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
    ECGB_REQUIRE(t->peer_area != nullptr && t->peer_world == world, "call ecgb_trainer_peer_area(world) first");
    ECGB_REQUIRE(d_areas[rank] == t->peer_area, "d_areas[rank] must be this trainer's own area");
    ECGB_REQUIRE(num_merges <= t->max_merges, "num_merges %u > max_merges %u", num_merges, t->max_merges);
    t->v.rank = rank;
    t->v.world = world;
    // the records of the initial stream (step 0 reads parity 0), as tagged units
    {
        std::vector<Boundary> all((size_t)world);
        ECGB_CUDA(cudaMemcpyAsync(all.data(), d_all_boundaries, sizeof(Boundary) * all.size(), cudaMemcpyDeviceToHost, st));
        ECGB_CUDA(cudaStreamSynchronize(st));
        std::vector<unsigned long long> units((size_t)world * 8, 0ull);
        const uint32_t tag = peer_tag(t->peer_epoch, 0);
        for (int r = 0; r < world; r++) {
            const Boundary &bd = all[r];
            unsigned long long *u = &units[(size_t)r * 8];
            u[0] = peer_unit(bd.n_lo, tag);
            u[1] = peer_unit(bd.n_hi, tag);
            u[2] = peer_unit((bd.first[0] & 0xFFFFu) | (bd.first[1] << 16), tag);
            u[3] = peer_unit((bd.first[2] & 0xFFFFu) | (bd.last[0] << 16), tag);
            u[4] = peer_unit(bd.last[1], tag);
        }
        ECGB_CUDA(cudaMemcpyAsync(t->peer_area + peer_rec_off(world, 0, 0), units.data(), units.size() * 8, cudaMemcpyHostToDevice, st));
        ECGB_CUDA(cudaStreamSynchronize(st));
    }
    TrainView view = t->v;
    view.redundant_max = kRedundantArgmax;
    const char *knob = getenv("ECGB_RESIDENT_TAIL");
    view.resident_ok = knob ? (uint32_t)atoi(knob) : 2u;
    PeerView pv{};
    pv.rank = rank;
    pv.world = world;
    pv.cap = t->peer_cap;
    pv.epoch = t->peer_epoch;
    for (int r = 0; r < world; r++) {
        ECGB_REQUIRE(d_areas[r] != nullptr, "d_areas[%d] is NULL", r);
        pv.area[r] = static_cast<uint8_t *>(d_areas[r]);
    }
    pv.out_count = t->peer_ctr;
    pv.done_ctas = t->peer_ctr + 2;
    pv.overflow = t->peer_ctr + 3;
    pv.timeout_ns = (unsigned long long)((timeout_s > 0 ? timeout_s : 30.0) * 1e9);
    *view_out = view;
    *pv_out = pv;
    return ECGB_OK;
}

template <class K>
static int dist_grid(const ecgb_trainer *t, K kernel, uint32_t max_ctas, int ranks_on_device, int *grid_out) {
    int per_sm = 0;
    const size_t dyn_smem = (size_t)kChunkCap * sizeof(uint16_t);
    ECGB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    ECGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTPB, dyn_smem));
    if (per_sm < 1) return fail(ECGB_EUNSUPPORTED, "the persistent sharded kernel does not fit on this device");
    int grid = std::min(t->sms * std::min(per_sm, kCtasPerSm), (int)t->v.cta_stride) / ranks_on_device;
    if (max_ctas > 0) grid = std::min(grid, (int)max_ctas);
    if (grid < 1) return fail(ECGB_EUNSUPPORTED, "%d ranks do not fit on one device", ranks_on_device);
    *grid_out = grid;
    return ECGB_OK;
}

// Launch the persistent sharded loop of this rank (asynchronous; every rank of the run must launch, each on
// its own device).  d_areas[r] = rank r's receive area as addressable from this process; d_all_boundaries =
// the gathered records of ecgb_trainer_dist_begin.  The table must hold the global histogram (dist_count +
// dist_apply).  Read the outcome with ecgb_trainer_results.
extern "C" int ecgb_trainer_dist_run(ecgb_trainer *t, int rank, int world, void *const *d_areas, const void *d_all_boundaries,
                                     uint32_t num_merges, uint32_t max_ctas, double timeout_s, void *stream) {
    ECGB_REQUIRE(t, "NULL argument");
    DeviceGuard g(t->device);
    cudaStream_t st = as_stream(stream);
    TrainView view;
    PeerView pv;
    int rc = dist_prepare(t, rank, world, d_areas, d_all_boundaries, num_merges, timeout_s, st, &view, &pv);
    if (rc) return rc;
    int grid = 0;
    rc = dist_grid(t, dist_loop_kernel, max_ctas, 1, &grid);
    if (rc) return rc;
    uint32_t steps = num_merges;
    void *kargs[] = {&view, &pv, &steps};
    ECGB_CUDA(cudaLaunchCooperativeKernel((const void *)dist_loop_kernel, dim3(grid), dim3(kTPB), kargs,
                                          (size_t)kChunkCap * sizeof(uint16_t), st));
    ECGB_CUDA(cudaGetLastError());
    t->device_steps = true;
    return ECGB_OK;
}

// The same run with every rank on ONE device, as one cooperative launch (blockIdx.y = rank) so that all ranks
// are co-resident by construction: ts[r] is rank r's trainer (all on the same device), d_areas[r] its area.
extern "C" int ecgb_trainer_dist_run_local(ecgb_trainer *const *ts, int world, void *const *d_areas, const void *d_all_boundaries,
                                           uint32_t num_merges, uint32_t max_ctas, double timeout_s, void *stream) {
    ECGB_REQUIRE(ts && world >= 1 && world <= kMaxWorld, "bad arguments");
    for (int r = 0; r < world; r++) ECGB_REQUIRE(ts[r] && ts[r]->device == ts[0]->device, "every trainer must live on the same device");
    DeviceGuard g(ts[0]->device);
    cudaStream_t st = as_stream(stream);
    std::vector<DistLaunch> h((size_t)world);
    for (int r = 0; r < world; r++) {
        int rc = dist_prepare(ts[r], r, world, d_areas, d_all_boundaries, num_merges, timeout_s, st, &h[r].v, &h[r].pv);
        if (rc) return rc;
    }
    int grid = 0;
    int rc = dist_grid(ts[0], dist_loop_local_kernel, max_ctas, world, &grid);
    if (rc) return rc;
    DistLaunch *d_ranks = nullptr;
    rc = dev_alloc(ts[0], (void **)&d_ranks, sizeof(DistLaunch) * (size_t)world, false);  // lives as long as rank 0's trainer
    if (rc) return rc;
    ECGB_CUDA(cudaMemcpyAsync(d_ranks, h.data(), sizeof(DistLaunch) * (size_t)world, cudaMemcpyHostToDevice, st));
    ECGB_CUDA(cudaStreamSynchronize(st));
    uint32_t steps = num_merges;
    const DistLaunch *arg0 = d_ranks;
    void *kargs[] = {&arg0, &steps};
    ECGB_CUDA(cudaLaunchCooperativeKernel((const void *)dist_loop_local_kernel, dim3(grid, world), dim3(kTPB), kargs,
                                          (size_t)kChunkCap * sizeof(uint16_t), st));
    ECGB_CUDA(cudaGetLastError());
    for (int r = 0; r < world; r++) ts[r]->device_steps = true;
    return ECGB_OK;
}

// Advance the device-side step counter (the last node of a captured step).
extern "C" int ecgb_trainer_dist_advance(ecgb_trainer *t, void *stream) {
    ECGB_REQUIRE(t, "NULL argument");
    DeviceGuard g(t->device);
    advance_step_kernel<<<1, 1, 0, as_stream(stream)>>>(t->v.dev);
    ECGB_CUDA(cudaGetLastError());
    t->device_steps = true;
    return ECGB_OK;
}

// Synchronise and read back the merges chosen so far (steps [0, n_steps)).
extern "C" int ecgb_trainer_results(ecgb_trainer *t, uint32_t n_steps, uint32_t *h_pairs, uint64_t *h_counts,
                                    uint32_t *h_ntied, uint32_t *n_done) {
    ECGB_REQUIRE(t && n_done, "NULL argument");
    ECGB_REQUIRE(n_steps <= t->max_merges, "n_steps out of range");
    DeviceGuard g(t->device);
    ECGB_CUDA(cudaDeviceSynchronize());
    int rc = check_tables(t);
    if (rc) return rc;
    DevState hs;
    ECGB_CUDA(cudaMemcpy(&hs, t->v.dev, sizeof(hs), cudaMemcpyDeviceToHost));
    if (t->peer_area != nullptr) print_phases();
    if (t->device_steps) {  // steps were driven by the device-side counter
        t->steps_done = std::min(hs.cur_step, t->max_merges);
        t->argmax_for = t->steps_done;
    }
    const uint32_t done = std::min(std::min(n_steps, hs.done_step), t->device_steps ? t->steps_done : n_steps);
    t->steps_done = std::min(t->steps_done, done);
    std::vector<Best> hb(done ? done : 1);
    if (done) ECGB_CUDA(cudaMemcpy(hb.data(), t->v.best, sizeof(Best) * done, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < done; i++) {
        if (h_pairs) { h_pairs[2 * i] = hb[i].key >> 16; h_pairs[2 * i + 1] = hb[i].key & 0xFFFFu; }
        if (h_counts) h_counts[i] = hb[i].count;
        if (h_ntied) h_ntied[i] = hb[i].ntied;
    }
    *n_done = done;
    return ECGB_OK;
}
