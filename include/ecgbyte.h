/*
 * ecgbyte.h -- C ABI of libecgbyte.so, the B200 (sm_100a) implementation of the
 * ECG-Byte tokenizer hot path.
 *
 * This is the drop-in boundary for the reference's only native component, the
 * PyO3 module `rust_bpe` (/root/reference/ecg_byte/rust_bpe/src/lib.rs:195-199)
 * and for the quantiser in front of it (ecg_byte/utils/tokenizer_utils.py:14-19).
 * Each entry point names the reference interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types in signatures.
 *     `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - pointers named d_* are DEVICE pointers on the handle's device, h_* are HOST
 *     pointers.  The library never frees caller memory.
 *   - every function returns an ecgb_status; ecgb_last_error() gives the message
 *     of the last failure on the calling thread.  Nothing throws or aborts across
 *     the ABI (the reference panics through .unwrap(), lib.rs:68,81,106-107).
 *   - handles are thread-compatible: distinct handles / streams may be used from
 *     distinct threads concurrently.  There is no global mutable state.
 *   - launches are asynchronous on `stream` unless the name ends in _host.
 *   - there is NO CPU fallback: every compute entry point fails with
 *     ECGB_ENODEVICE when no CUDA device is usable.
 */
#ifndef ECGBYTE_H
#define ECGBYTE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    ECGB_OK = 0,
    ECGB_EINVAL = 1,    /* bad argument (message says which) */
    ECGB_ENOMEM = 2,    /* host or device allocation failed */
    ECGB_ECUDA = 3,     /* CUDA runtime error (message carries cudaGetErrorString) */
    ECGB_ENODEVICE = 4, /* no usable CUDA device */
    ECGB_ECAPACITY = 5, /* an output buffer / table capacity was too small */
    ECGB_EUNSUPPORTED = 6
} ecgb_status;

/* stored sample types of a record (float64 is the reference's
 * on-disk type, preprocess_utils.py:94,220-223; fp32 / int16 are the compact
 * variants named by BASELINE.json).  ECGB_U8 = already-quantised text bytes. */
typedef enum { ECGB_F32 = 0, ECGB_F64 = 1, ECGB_I16 = 2, ECGB_U8 = 3 } ecgb_dtype;

typedef struct ecgb_quantizer ecgb_quantizer;
typedef struct ecgb_vocab ecgb_vocab;
typedef struct ecgb_trainer ecgb_trainer;

const char *ecgb_last_error(void);
int ecgb_version(void);
/* number of CUDA devices visible (0 and ECGB_ENODEVICE when there is none) */
int ecgb_device_count(int *n_out);

/* ------------------------------------------------------------------------- */
/* Q1  normalize_all  (tokenizer_utils.py:14-19; ALPHABET tokenizer_utils.py:12) */
/* ------------------------------------------------------------------------- */
/* symbol = 'a' + min(floor(clip((s - (p1-0.5)) / ((p99+0.5)-(p1-0.5)+1e-6), 0, 1) * 26), 25),
 * evaluated in float64 in exactly that order.  The expression is a non-decreasing
 * step function of s, so it is fully described by 25 thresholds t_k = the smallest
 * representable sample with symbol >= k; the quantiser object holds them (found by
 * bisection with the float64 expression itself) and the kernels classify against
 * them -- bit-identical to the expression, without a float64 divide per sample.
 * int16 samples are de-scaled first: s = (double)v * i16_scale.
 * NaN -> 'a' (NumPy's NaN->uint8 cast on x86-64). Fails with ECGB_EINVAL when the
 * denominator is not > 0 or a percentile is not finite. */
int ecgb_quantizer_create(double p1, double p99, ecgb_dtype dtype, double i16_scale, int device,
                          ecgb_quantizer **out);
int ecgb_quantizer_destroy(ecgb_quantizer *q);
/* the 25 thresholds as doubles (every stored type converts to double exactly) */
int ecgb_quantizer_thresholds(const ecgb_quantizer *q, double h_thr_out[25]);
/* d_in: n samples of the quantiser's dtype -> d_out: n bytes 'a'..'z' */
int ecgb_quantize(const ecgb_quantizer *q, const void *d_in, size_t n, uint8_t *d_out, void *stream);
/* same result computed sample by sample with the float64 expression on the device
 * (reference operation order, IEEE divide, no FMA) -- the on-device cross-check */
int ecgb_quantize_direct(const ecgb_quantizer *q, const void *d_in, size_t n, uint8_t *d_out,
                         void *stream);
/* host buffers in/out (H2D, kernel, D2H, synchronous) */
int ecgb_quantize_host(const ecgb_quantizer *q, const void *h_in, size_t n, uint8_t *h_out);

/* ------------------------------------------------------------------------- */
/* E1  TrieNode / trie build  (lib.rs:127-147, 153-161)                      */
/* ------------------------------------------------------------------------- */
/* merges in the reference's pickle form, flattened: merge i has the expanded
 * base-symbol sequence seq[seq_off[i] .. seq_off[i+1]) and token id ids[i]
 * (list[tuple[list[int], int]], lib.rs:110,150).  All 256 single bytes are
 * inserted first (lib.rs:155-157); a later duplicate sequence overwrites the
 * token id (lib.rs:145).  The trie is flattened once and kept resident on
 * `device` (the reference rebuilds it on every encode call). */
int ecgb_vocab_create(const uint32_t *h_seq, const uint64_t *h_seq_off, const uint32_t *h_ids,
                      uint32_t n_merges, int device, ecgb_vocab **out);
int ecgb_vocab_destroy(ecgb_vocab *v);

typedef struct {
    uint32_t n_merges;
    uint32_t n_nodes;       /* trie nodes incl. root */
    uint32_t n_classes;     /* distinct symbols that occur inside merges (+ a..z) */
    uint32_t compact;       /* 1: 8-byte bitmap nodes (<= 31 classes), 0: wide nodes */
    uint32_t max_token_len; /* longest expanded sequence */
    uint32_t node_bytes;    /* size of the device node table */
    uint32_t smem_nodes;    /* nodes the bitmap-trie kernels keep in shared memory */
    uint32_t pair_slots;    /* slots of the two-symbol-stride table the fused encoder walks (0: none) */
} ecgb_vocab_info_t;
int ecgb_vocab_info(const ecgb_vocab *v, ecgb_vocab_info_t *out);

/* Host-only view of the flattened trie the fused encoder walks (inspection and tests; no device
 * needed): the two-symbol-stride "pair table" of the same merges (lib.rs:153-161 trie, states =
 * nodes at even depth; layout documented in csrc/trie_host.h).  Two-call sizing through *n_ent_out
 * (h_ent / h_tok may be NULL).  meta = {root_base, dead_base, W, NC, SM, SE, states, slots used};
 * h_cls_out[b] = class of byte b (255: none). */
int ecgb_pairtab_host(const uint32_t *h_seq, const uint64_t *h_seq_off, const uint32_t *h_ids,
                      uint32_t n_merges, uint32_t *h_ent, uint16_t *h_tok, uint32_t cap,
                      uint32_t *n_ent_out, uint32_t meta[8], uint8_t h_cls_out[256]);

/* ------------------------------------------------------------------------- */
/* E2  encode_text  (lib.rs:149-193) -- greedy longest match over the trie    */
/* ------------------------------------------------------------------------- */
/* Batch of records of text bytes.  Record r is d_sym[off[r] .. off[r+1]) when
 * d_offsets != NULL (n_rec+1 entries), else d_sym[r*rec_len .. (r+1)*rec_len).
 * Tokens of record r go to d_tokens[r*out_stride ..]; d_len[r] is the TRUE token
 * count (tokens beyond out_stride are counted but not stored). */
int ecgb_encode_symbols(const ecgb_vocab *v, const uint8_t *d_sym, size_t n_rec, size_t rec_len,
                        const uint64_t *d_offsets, int32_t *d_tokens, size_t out_stride,
                        int32_t *d_len, void *stream);
/* Fused Q1+E2 (data_loader.py:74-76): raw samples -> token ids; record r is the
 * rec_len = C*L samples at d_in + r*rec_len, lead-major (C-order flatten,
 * tokenizer_utils.py:59).  Symbols never touch HBM.  The vocabulary must map
 * 'a'..'z' (every ECG-Byte vocabulary does). */
int ecgb_encode_batch(const ecgb_vocab *v, const ecgb_quantizer *q, const void *d_in, size_t n_rec,
                      size_t rec_len, int32_t *d_tokens, size_t out_stride, int32_t *d_len,
                      void *stream);
/* rust_bpe.encode_text(text, merges) for one string with host buffers: returns
 * the token count in *n_out; ECGB_ECAPACITY (with *n_out set) when cap is too small. */
int ecgb_encode_text_host(const ecgb_vocab *v, const uint8_t *h_text, size_t n, uint32_t *h_out,
                          size_t cap, size_t *n_out);
/* Same as ecgb_encode_batch with HOST sample / token buffers (pinned or pageable):
 * H2D copy, kernel, D2H copy of tokens and lengths, synchronous. */
int ecgb_encode_batch_host(const ecgb_vocab *v, const ecgb_quantizer *q, const void *h_in,
                           size_t n_rec, size_t rec_len, int32_t *h_tokens, size_t out_stride,
                           int32_t *h_len);

/* ------------------------------------------------------------------------- */
/* Either side of the encoder                                  */
/* ------------------------------------------------------------------------- */
/* decode_text (tokenizer_utils.py:75-77): every token of record r (d_tokens[r*in_stride ..],
 * d_len[r] of them) is replaced by its expanded bytes; d_sym_len[r] = bytes produced
 * (bytes beyond sym_stride are counted, not stored).  Synchronises the stream; an id that is
 * not in the vocabulary gives ECGB_EINVAL (the reference raises KeyError). */
int ecgb_decode_symbols(const ecgb_vocab *v, const int32_t *d_tokens, size_t n_rec, size_t in_stride,
                        const int32_t *d_len, uint8_t *d_sym, size_t sym_stride, int32_t *d_sym_len,
                        void *stream);
/* expand_attention (runners/interpret.py:106-111): one attention value per token -> one per base
 * symbol (the value of token i repeated len(token i) times), per record; same layout rules as
 * ecgb_decode_symbols.  (Lengths are counted in symbols; the reference counts characters of the
 * vocab string, which differs only for raw bytes > 127 spelled "<b>".) */
int ecgb_expand_attention(const ecgb_vocab *v, const int32_t *d_tokens, const float *d_attn, size_t n_rec,
                          size_t in_stride, const int32_t *d_len, float *d_out, size_t out_stride,
                          int32_t *d_out_len, void *stream);
/* Compact (CSR) copy of the encoder's output for host consumers of ecgb_encode_batch (the list[int] that
 * rust_bpe.encode_text returns, lib.rs:192, per record): 2-byte ids (ids < 65 536), rows back to back.
 * d_off[n_rec + 1] = row offsets in tokens counted from `base`, d_off[n_rec] = base + total; row r is written to
 * out[d_off[r] ...].  `out` may be pinned (mapped) host memory: the kernel then stores straight over PCIe. */
int ecgb_tokens_csr(const int32_t *d_tokens, size_t in_stride, const int32_t *d_len, size_t n_rec, uint16_t *out,
                    uint64_t *d_off, uint64_t base, int device, void *stream);
/* analyze_token_distribution (tokenizer_utils.py:30-54): Counter over the encoded ids of a batch.
 * d_counts[n_ids] (u64, device) is ACCUMULATED into -- zero it first; the per-record token_lengths
 * of the reference are the encoder's d_len.  An id outside [0, n_ids) is ECGB_EINVAL. */
int ecgb_token_histogram(const int32_t *d_tokens, size_t in_stride, const int32_t *d_len, size_t n_rec,
                         uint32_t n_ids, unsigned long long *d_counts, int device, void *stream);
/* reverse_normalize_all (tokenizer_utils.py:22-28): value = (symbol_index / 25) * ((p99+0.5) -
 * (p1-0.5)) + (p1-0.5) in float64, in that order (note 25 = len(ALPHABET) - 1, as the reference) */
int ecgb_dequantize(double p1, double p99, const uint8_t *d_sym, size_t n, double *d_out, int device,
                    void *stream);

/* compute_global_stats (preprocess_utils.py:168-213), the step that produces the percentiles
 * the quantiser consumes: np.min / np.max over every stored sample (NaN propagates) ... */
int ecgb_minmax(const void *d_in, ecgb_dtype dtype, size_t n, double *h_min, double *h_max, int device,
                void *stream);
/* ... and np.percentile(samples, q) with NumPy's default 'linear' method (preprocess_utils.py:205-206)
 * for nq percentiles at once; d_samples are float64 on the device.  Synchronises the stream. */
int ecgb_percentiles(const double *d_samples, size_t n, const double *h_q, int nq, double *h_out, int device,
                     void *stream);

/* ECGTokenDataset post-processing (data_loader.py:80, 26-31, 101-132), one row per sample:
 *   row = [pad]*k + [bos, sig_start] + lut[signal tokens][:available] + [sig_end] + question +
 *         answer + [eos],  available = pad_to_max - len(question) - len(answer),
 *   row length pad_to_max + 4; labels are -100 up to the end of the question; attention mask
 *   0 on pad; position ids = cumsum(mask) - 1 (0 on pad).
 * d_lut[k] is the LLM id of the string 'signal_k' (main.py:144-146); question+answer ids of
 * sample r are d_text[d_text_off[r] .. d_text_off[r+1]) with the first d_q_len[r] the question.
 * d_status[r] = 1 when question+answer exceed pad_to_max (the reference's assert fails there). */
typedef struct {
    int64_t pad_id, bos_id, eos_id, sig_start_id, sig_end_id;
    uint32_t pad_to_max;
} ecgb_pack_cfg;
int ecgb_pack_training(const int32_t *d_tokens, size_t in_stride, const int32_t *d_len, size_t n_rec,
                       const int64_t *d_lut, uint32_t lut_size, const int64_t *d_text,
                       const uint64_t *d_text_off, const int32_t *d_q_len, const ecgb_pack_cfg *cfg,
                       int64_t *d_input_ids, float *d_attn_mask, int64_t *d_labels,
                       int64_t *d_position_ids, int32_t *d_status, int device, void *stream);

/* ------------------------------------------------------------------------- */
/* T1-T4  byte_pair_encoding  (lib.rs:58-125)                                */
/* ------------------------------------------------------------------------- */
/* The corpus is ONE string (tokenizer_utils.py:93): pairs are counted and merged
 * across record boundaries.  A trainer owns a contiguous shard of that string
 * (the whole string when world_size == 1).
 *
 * Tie rule (the reference's winner depends on hash-map iteration order,
 * lib.rs:92-94): maximum count, then the lexicographically smallest (left,right).
 * Counts are 64-bit (the reference's u32 counts wrap). */
/* table_log2: log2 of the pair-histogram capacity in slots (0 = default 2^22); the
 * run fails with ECGB_ECAPACITY (never silently) if the distinct pairs outgrow it. */
int ecgb_trainer_create(int device, uint64_t capacity_tokens, uint32_t max_merges, uint32_t table_log2,
                        ecgb_trainer **out);
int ecgb_trainer_destroy(ecgb_trainer *t);
/* load this rank's shard: n bytes of text (device or host pointer) */
int ecgb_trainer_load_device(ecgb_trainer *t, const uint8_t *d_text, uint64_t n, void *stream);
int ecgb_trainer_load_host(ecgb_trainer *t, const uint8_t *h_text, uint64_t n);
/* Single-device training loop: runs up to num_merges merge steps entirely on the
 * device (count -> argmax -> merge-apply, lib.rs:85-117) and returns the number
 * done (early stop when no pair is left, lib.rs:88-90).
 * h_pairs[2*i], h_pairs[2*i+1] = (left,right) of merge i (new id 256+i);
 * h_counts[i] = its count; h_ntied[i] = number of pairs sharing that count. */
int ecgb_trainer_run(ecgb_trainer *t, uint32_t num_merges, uint32_t *h_pairs, uint64_t *h_counts,
                     uint32_t *h_ntied, uint32_t *n_done);
/* current length of the merged token stream, and a copy of it (lib.rs:124 `ids`) */
int ecgb_trainer_length(ecgb_trainer *t, uint64_t *n_out);
int ecgb_trainer_ids_host(ecgb_trainer *t, uint32_t *h_ids, uint64_t cap, uint64_t *n_out);

/* h_n[i] = length of this shard's token stream after i merge steps, i in [0, n_steps] */
int ecgb_trainer_lengths(ecgb_trainer *t, uint32_t n_steps, uint64_t *h_n);
/* track_encoding (tokenizer_utils.py:95-134) with pair-form merges: apply the given (left, right) -> new id merges
 * in order to the loaded text (merge, lib.rs:10-26, once per pair); result through ecgb_trainer_ids_host. */
int ecgb_trainer_apply_pairs(ecgb_trainer *t, const uint32_t *h_pairs, const uint32_t *h_new_ids, uint32_t n);
/* pair-table occupancy: h_out = {slots claimed, capacity, argmax candidates listed, overflow flag} */
int ecgb_trainer_table_stats(ecgb_trainer *t, uint64_t h_out[4]);
/* every pair with a non-zero count in the live histogram (== get_stats, lib.rs:28-48, of
 * the current stream); two-call sizing through *n_out / ECGB_ECAPACITY */
int ecgb_trainer_histogram(ecgb_trainer *t, uint32_t *h_pairs, int64_t *h_counts, uint64_t cap,
                           uint64_t *n_out);

/* Step-wise interface for a corpus sharded over ranks (one trainer per rank holds a
 * CONTIGUOUS piece of the one corpus string, in rank order).  All calls
 * are asynchronous on `stream`; the two exchanges per step are plain all-gathers of
 * fixed-size device buffers done by the host (torch.distributed / NCCL):
 *
 *   dist_begin(rank, world)            -> boundary record            | all-gather
 *   dist_count(all boundaries)         -> delta list (local get_stats)| all-gather
 *   for step in 0..M:
 *     dist_commit(step, all lists)     -> applies every rank's list to the local copy
 *                                         of the GLOBAL histogram, argmax (identical on
 *                                         every rank), boundary record | all-gather
 *     dist_merge(step, all boundaries) -> merges in this shard (halo tokens and run
 *                                         parity come from the records), delta list
 *                                                                     | all-gather
 *   results()                          -> merges, counts, tie log
 *
 * Every rank applies the same lists in the same order, so the histograms -- and the
 * argmax -- are identical without any reduction.
 *
 * step = ECGB_STEP_DEVICE in dist_commit / dist_merge: take the step number from a counter on the
 * device, which ecgb_trainer_dist_advance increments.  The calls of one step then have constant
 * arguments, so commit / all-gather / merge / all-gather / advance can be captured once in a CUDA
 * graph and replayed per merge (ecgbyte/dist_train.py); replays beyond max_merges do nothing. */
#define ECGB_STEP_DEVICE 0xFFFFFFFFu
int ecgb_trainer_dist_sizes(const ecgb_trainer *t, uint32_t *boundary_bytes, uint32_t *list_bytes);
int ecgb_trainer_dist_begin(ecgb_trainer *t, int rank, int world, void *d_boundary_out, void *stream);
int ecgb_trainer_dist_count(ecgb_trainer *t, const void *d_all_boundaries, void *d_list_out, void *stream);
int ecgb_trainer_dist_commit(ecgb_trainer *t, uint32_t step, const void *d_all_lists, void *d_boundary_out,
                             void *stream);
int ecgb_trainer_dist_merge(ecgb_trainer *t, uint32_t step, const void *d_all_boundaries, void *d_list_out,
                            void *stream);
int ecgb_trainer_dist_advance(ecgb_trainer *t, void *stream);
int ecgb_trainer_results(ecgb_trainer *t, uint32_t n_steps, uint32_t *h_pairs, uint64_t *h_counts,
                         uint32_t *h_ntied, uint32_t *n_done);

/* Persistent sharded loop (lib.rs:85-117 over a corpus cut into contiguous shards): ONE
 * cooperative kernel per rank runs every merge step; the per-step exchange is device-initiated --
 * each rank writes its histogram patches and its 64-byte shard record straight into a receive area in
 * every peer's memory (NVLink peer stores + flag words), no launches, host calls or collectives inside
 * the loop.  Set-up (host; the step-wise calls above provide the initial records and histogram):
 *
 *   peer_area(world)            -> this rank's receive area             | exchange addresses (IPC handles
 *                                                                         between processes), BARRIER
 *   dist_begin / dist_count     -> record, local get_stats              | all-gather each (as above)
 *   dist_apply(all lists)       -> global histogram on every rank
 *   dist_run(rank, world, areas, all boundaries, M)   -- asynchronous; every rank must launch
 *   results()
 *
 * A rank whose peer never shows up gives up after timeout_s (default 30 s) and results() reports it.
 * max_ctas > 0 caps the grid. */
int ecgb_trainer_peer_area(ecgb_trainer *t, int world, void **d_area, uint64_t *bytes);
int ecgb_trainer_dist_apply(ecgb_trainer *t, const void *d_all_lists, int world, void *stream);
int ecgb_trainer_dist_run(ecgb_trainer *t, int rank, int world, void *const *d_areas, const void *d_all_boundaries,
                          uint32_t num_merges, uint32_t max_ctas, double timeout_s, void *stream);
/* every rank on ONE device as one cooperative launch (co-residency by construction): ts[r] = rank r's trainer */
int ecgb_trainer_dist_run_local(ecgb_trainer *const *ts, int world, void *const *d_areas, const void *d_all_boundaries,
                                uint32_t num_merges, uint32_t max_ctas, double timeout_s, void *stream);
/* CUDA IPC for the areas of ranks that live in other processes (one process per GPU) */
int ecgb_ipc_export(const void *d_ptr, uint8_t handle[64]);
int ecgb_ipc_open(const uint8_t handle[64], int device, void **d_ptr);
int ecgb_ipc_close(void *d_ptr, int device);

/* expanded sequences of merges given as pairs (lib.rs:101-110); two-call sizing:
 * h_seq_off[n_merges] always receives the total length. */
int ecgb_expand_merges(const uint32_t *h_pairs, uint32_t n_merges, uint32_t *h_seq, uint64_t seq_cap,
                       uint64_t *h_seq_off);

#ifdef __cplusplus
}
#endif
#endif /* ECGBYTE_H */
