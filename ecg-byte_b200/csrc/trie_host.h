// Host-side trie of a merges table (reference: rust_bpe/src/lib.rs:127-147 TrieNode,
// 153-161 build) and the two flattened forms the device kernels walk.
#pragma once
#include <cstdint>
#include <map>
#include <vector>

namespace ecgb {

struct HostNode {
    std::map<uint32_t, int> child;  // ordered by symbol
    int64_t token = -1;
};

struct HostTrie {
    std::vector<HostNode> nodes;
    uint32_t max_len = 1;
    int insert(const uint32_t *seq, size_t len, uint32_t id) {
        int n = 0;
        for (size_t i = 0; i < len; i++) {
            auto it = nodes[n].child.find(seq[i]);
            if (it == nodes[n].child.end()) {
                nodes.emplace_back();
                int c = (int)nodes.size() - 1;
                nodes[n].child[seq[i]] = c;
                n = c;
            } else {
                n = it->second;
            }
        }
        nodes[n].token = id;  // lib.rs:145: later insert overwrites
        return n;
    }
};

// lib.rs:155-161: all 256 single bytes, then every merge in list order.  Returns 0 or the
// (1-based) index of the first malformed merge.
int build_host_trie(HostTrie *t, const uint32_t *seq, const uint64_t *seq_off, const uint32_t *ids, uint32_t n_merges);

// Symbol classes of a vocabulary: 'a'..'z' -> 0..25 (always), then every other byte that needs a
// trie edge (it occurs inside a merge, or its single-byte token id was overwritten) in byte order.
// cls[b] = 255 for bytes without a class: such a byte is its own token and ends every walk.
void vocab_classes(const HostTrie &t, uint8_t cls[256], int *n_classes);

// Two-symbol-stride double-array automaton over the trie ("pair table"), the structure the
// fused encoder walks.  A STATE is a trie node at an even depth below the start of a token.
// From state s the next two symbol classes (a, b) select slot  base[s] + (a << W | b):
//   ent[slot] = next_base << 16 | code << 2 | T2 << 1 | T1      (bits 14-15 are zero)
//     code  = a << W | b : the slot belongs to s iff the stored code equals the probing one
//             (every state has a distinct base, so base + code identifies (s, code));
//     T1    = the node after a is a token, T2 = the node after (a, b) is a token;
//     next_base = base of the node after (a, b) (a base that matches nothing if it is a leaf).
//   A token that ends after an odd number of symbols is found through the "single" slot
//   base[s] + (a << W | SM): it exists iff the node after a is a token.
//   tok[slot] = token id of the node the slot leads to (pair: after (a, b); single: after a).
// Classes: 0..NC-1 symbols, SM = NC (single marker), SE = NC + 1 (sentinel: end of record, or
// a byte without a class); 2^W - 1 is never used, so an empty slot (all ones) matches nothing.
struct PairTab {
    std::vector<uint32_t> ent;
    std::vector<uint16_t> tok;
    uint32_t root_base = 0, dead_base = 0;
    uint32_t W = 5, NC = 0, SM = 0, SE = 0;
    uint32_t n_states = 0, n_slots_used = 0;
};

// cls: byte -> class (>= n_classes: no class).  false when the vocabulary does not fit the
// format (too many classes, token ids >= 2^16, more than 65535 slots).
bool build_pairtab(const HostTrie &t, const uint8_t cls[256], int n_classes, PairTab *out);

}  // namespace ecgb
