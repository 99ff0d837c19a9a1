// E1: trie construction (reference: rust_bpe/src/lib.rs:127-147 TrieNode, 153-161 build).
//
// The reference rebuilds a HashMap-of-HashMaps trie on every encode_text call.  Here
// the trie is built once per vocabulary on the host, renumbered breadth-first so the
// children of a node are contiguous, and flattened into bitmap nodes that stay
// resident on the device:
//   compact node (8 bytes, alphabets of <= 31 symbol classes -- every ECG-Byte
//   vocabulary: 26 letters):  x = child bitmap over classes,
//                             y = first_child << 16 | (token_id + 1)  (0 = not a token)
//   child(c) = first_child + popc(bitmap & ((1 << c) - 1))          -- no hashing, one
//   8-byte shared-memory load per trie step.
//   wide node (40 bytes, any byte alphabet): 256-bit bitmap, first_child, token_id.
#include <algorithm>
#include <cstring>
#include <map>
#include <new>
#include <vector>

#include "common.h"
#include "trie_host.h"

struct ecgb_vocab {
    int device = 0;
    ecgb_vocab_info_t info{};
    ecgb::VocabView view{};
    void *d_nodes = nullptr;
    uint8_t *d_cls = nullptr;
    uint8_t *d_dec_sym = nullptr;
    uint32_t *d_dec_off = nullptr;
    uint32_t *d_pair_ent = nullptr;  // pair table (trie_host.h), NULL when the vocabulary does not fit it
    uint16_t *d_pair_tok = nullptr;
    uint8_t *d_pair_cls = nullptr;
    uint8_t h_cls[256];
    // host copy of the expanded sequences (decode, pickles)
    std::vector<uint32_t> seq;
    std::vector<uint64_t> seq_off;
    std::vector<uint32_t> ids;
};

const ecgb::VocabView *ecgb_vocab_view(const ecgb_vocab *v) { return &v->view; }
int ecgb_vocab_device(const ecgb_vocab *v) { return v->device; }

using namespace ecgb;

extern "C" int ecgb_vocab_create(const uint32_t *h_seq, const uint64_t *h_seq_off, const uint32_t *h_ids,
                                 uint32_t n_merges, int device, ecgb_vocab **out) {
    ECGB_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    ECGB_REQUIRE(n_merges == 0 || (h_seq && h_seq_off && h_ids), "NULL merge arrays");
    int rc = check_device(device);
    if (rc) return rc;

    HostTrie t;
    {
        const int bad = build_host_trie(&t, h_seq, h_seq_off, h_ids, n_merges);
        ECGB_REQUIRE(bad == 0, "merge %d is malformed (empty sequence, or an element that is not a byte)", bad - 1);
    }

    // ---- symbol classes: a..z -> 0..25, other bytes that need a trie edge -> 26.. ----
    uint8_t cls_all[256];
    int n_classes = 0;
    vocab_classes(t, cls_all, &n_classes);
    // the bitmap nodes hold up to 31 classes (class 31 = none)
    bool compact = n_classes <= 31;
    uint8_t cls[256];
    for (int b = 0; b < 256; b++) cls[b] = cls_all[b] < 31 ? cls_all[b] : (uint8_t)31;
    int64_t max_tok = 255;
    for (auto &nd : t.nodes) max_tok = std::max(max_tok, nd.token);
    if (max_tok >= 0xFFFF) compact = false;

    ecgb_vocab *v = new (std::nothrow) ecgb_vocab();
    if (!v) return fail(ECGB_ENOMEM, "host allocation failed");
    v->device = device;

    // ---- breadth-first renumbering ----
    std::vector<int> order;  // new index -> old index
    std::vector<uint32_t> first_child;
    order.reserve(t.nodes.size());
    order.push_back(0);
    size_t n_total = 0;
    std::vector<uint32_t> words;
    for (int pass = 0; pass < 2; pass++) {
        // pass 0: try compact (may overflow 16-bit indices); pass 1: wide
        if (pass == 1) compact = false;
        order.assign(1, 0);
        first_child.clear();
        for (size_t head = 0; head < order.size(); head++) {
            const HostNode &nd = t.nodes[order[head]];
            first_child.push_back((uint32_t)order.size());
            if (compact) {
                // children ordered by class; bytes without a class are handled at the root only
                std::vector<std::pair<int, int>> ch;
                for (auto &kv : nd.child)
                    if (cls[kv.first] < 31) ch.push_back({cls[kv.first], kv.second});
                std::sort(ch.begin(), ch.end());
                for (auto &c : ch) order.push_back(c.second);
            } else {
                for (auto &kv : nd.child) order.push_back(kv.second);
            }
        }
        n_total = order.size();
        if (compact && n_total > 0xFFFF) continue;  // retry wide
        break;
    }

    if (compact) {
        words.resize(n_total * 2);
        for (size_t i = 0; i < n_total; i++) {
            const HostNode &nd = t.nodes[order[i]];
            uint32_t mask = 0;
            for (auto &kv : nd.child)
                if (cls[kv.first] < 31) mask |= 1u << cls[kv.first];
            uint32_t tok = nd.token >= 0 ? (uint32_t)nd.token + 1u : 0u;
            words[2 * i] = mask;
            words[2 * i + 1] = (first_child[i] << 16) | tok;
        }
    } else {
        words.resize(n_total * 10);
        for (size_t i = 0; i < n_total; i++) {
            const HostNode &nd = t.nodes[order[i]];
            uint32_t *w = &words[10 * i];
            std::memset(w, 0, 40);
            for (auto &kv : nd.child) w[kv.first >> 5] |= 1u << (kv.first & 31);
            w[8] = first_child[i];
            w[9] = nd.token >= 0 ? (uint32_t)nd.token : 0xFFFFFFFFu;
        }
        for (int b = 0; b < 256; b++) cls[b] = (uint8_t)b;  // unused by the wide kernel
    }

    v->info.n_merges = n_merges;
    v->info.n_nodes = (uint32_t)n_total;
    v->info.n_classes = (uint32_t)n_classes;
    v->info.compact = compact ? 1u : 0u;
    v->info.max_token_len = t.max_len;
    v->info.node_bytes = (uint32_t)(words.size() * 4);
    std::memcpy(v->h_cls, cls, 256);
    if (n_merges) {
        v->seq.assign(h_seq, h_seq + h_seq_off[n_merges]);
        v->seq_off.assign(h_seq_off, h_seq_off + n_merges + 1);
        v->ids.assign(h_ids, h_ids + n_merges);
    }

    DeviceGuard g(device);
    cudaError_t e = cudaMalloc(&v->d_nodes, words.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&v->d_cls, 256);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_nodes, words.data(), words.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_cls, cls, 256, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(v->d_nodes);
        cudaFree(v->d_cls);
        delete v;
        return fail(e == cudaErrorMemoryAllocation ? ECGB_ENOMEM : ECGB_ECUDA, "vocab upload failed: %s", cudaGetErrorString(e));
    }
    // shared-memory residency: leave room for the quantiser tables and bookkeeping
    int smem_max = 0;
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    size_t budget = smem_max > 24 * 1024 ? (size_t)smem_max - 16 * 1024 : 0;
    uint32_t fit = (uint32_t)std::min<size_t>(n_total, budget / 8);
    v->info.smem_nodes = compact ? fit : 0;

    v->view.d_nodes = static_cast<const uint2 *>(v->d_nodes);
    v->view.d_wide = static_cast<const uint32_t *>(v->d_nodes);
    v->view.n_nodes = (uint32_t)n_total;
    v->view.smem_nodes = v->info.smem_nodes;
    v->view.d_cls = v->d_cls;
    v->view.compact = compact ? 1 : 0;
    v->view.ecg_alphabet = 1;  // classes 0..25 are always 'a'..'z' in the compact layout
    v->view.max_token_len = t.max_len;

    // ---- pair table: the two-symbol-stride automaton the fused encoder walks (trie_host.h) ----
    {
        PairTab pt;
        if (build_pairtab(t, cls_all, n_classes, &pt)) {
            uint8_t pcls[256];
            for (int b = 0; b < 256; b++) pcls[b] = cls_all[b] < n_classes ? cls_all[b] : (uint8_t)pt.SE;
            cudaError_t e4 = cudaMalloc((void **)&v->d_pair_ent, pt.ent.size() * 4);
            if (e4 == cudaSuccess) e4 = cudaMalloc((void **)&v->d_pair_tok, pt.tok.size() * 2 + 16);
            if (e4 == cudaSuccess) e4 = cudaMalloc((void **)&v->d_pair_cls, 256);
            if (e4 == cudaSuccess) e4 = cudaMemcpy(v->d_pair_ent, pt.ent.data(), pt.ent.size() * 4, cudaMemcpyHostToDevice);
            if (e4 == cudaSuccess) e4 = cudaMemcpy(v->d_pair_tok, pt.tok.data(), pt.tok.size() * 2, cudaMemcpyHostToDevice);
            if (e4 == cudaSuccess) e4 = cudaMemcpy(v->d_pair_cls, pcls, 256, cudaMemcpyHostToDevice);
            if (e4 != cudaSuccess) {
                ecgb_vocab_destroy(v);
                return fail(e4 == cudaErrorMemoryAllocation ? ECGB_ENOMEM : ECGB_ECUDA, "pair table upload failed: %s", cudaGetErrorString(e4));
            }
            v->view.pair.d_ent = v->d_pair_ent;
            v->view.pair.d_tok = v->d_pair_tok;
            v->view.pair.d_cls = v->d_pair_cls;
            v->view.pair.n_ent = (uint32_t)pt.ent.size();
            v->view.pair.root_base = pt.root_base;
            v->view.pair.W = pt.W;
            v->view.pair.SM = pt.SM;
            v->view.pair.SE = pt.SE;
            v->info.pair_slots = (uint32_t)pt.ent.size();
        }
    }

    // ---- decode tables: id -> expanded bytes (a later merge with the same id wins, like a dict) ----
    {
        uint32_t max_id = 255;
        for (uint32_t i = 0; i < n_merges; i++) max_id = std::max(max_id, h_ids[i]);
        if (max_id < (1u << 24)) {
            std::vector<int64_t> which((size_t)max_id + 1, -1);
            for (uint32_t i = 0; i < n_merges; i++) which[h_ids[i]] = (int64_t)i;
            std::vector<uint32_t> off((size_t)max_id + 2, 0);
            std::vector<uint8_t> bytes;
            for (uint32_t id = 0; id <= max_id; id++) {
                off[id] = (uint32_t)bytes.size();
                if (which[id] >= 0) {
                    const uint64_t o = h_seq_off[which[id]], e2 = h_seq_off[which[id] + 1];
                    for (uint64_t k = o; k < e2; k++) bytes.push_back((uint8_t)h_seq[k]);
                } else if (id < 256) {
                    bytes.push_back((uint8_t)id);
                }
            }
            off[(size_t)max_id + 1] = (uint32_t)bytes.size();
            cudaError_t e3 = cudaMalloc((void **)&v->d_dec_sym, bytes.size() + 16);
            if (e3 == cudaSuccess) e3 = cudaMalloc((void **)&v->d_dec_off, off.size() * 4);
            if (e3 == cudaSuccess) e3 = cudaMemcpy(v->d_dec_sym, bytes.data(), bytes.size(), cudaMemcpyHostToDevice);
            if (e3 == cudaSuccess) e3 = cudaMemcpy(v->d_dec_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice);
            if (e3 != cudaSuccess) {
                ecgb_vocab_destroy(v);
                return fail(ECGB_ECUDA, "decode table upload failed: %s", cudaGetErrorString(e3));
            }
            v->view.d_dec_sym = v->d_dec_sym;
            v->view.d_dec_off = v->d_dec_off;
            v->view.dec_ids = max_id + 1;
        }
    }
    *out = v;
    return ECGB_OK;
}

extern "C" int ecgb_vocab_destroy(ecgb_vocab *v) {
    if (!v) return ECGB_OK;
    {
        DeviceGuard g(v->device);
        cudaFree(v->d_nodes);
        cudaFree(v->d_cls);
        cudaFree(v->d_dec_sym);
        cudaFree(v->d_dec_off);
        cudaFree(v->d_pair_ent);
        cudaFree(v->d_pair_tok);
        cudaFree(v->d_pair_cls);
    }
    delete v;
    return ECGB_OK;
}

extern "C" int ecgb_vocab_info(const ecgb_vocab *v, ecgb_vocab_info_t *out) {
    ECGB_REQUIRE(v && out, "NULL argument");
    *out = v->info;
    return ECGB_OK;
}
