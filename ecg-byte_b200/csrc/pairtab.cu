// E1 (lib.rs:127-161) flattened for the two-symbol-stride encoder: see PairTab in trie_host.h.
// Host code only.
#include <algorithm>
#include <cstring>

#include "common.h"
#include "trie_host.h"

namespace ecgb {

int build_host_trie(HostTrie *t, const uint32_t *seq, const uint64_t *seq_off, const uint32_t *ids, uint32_t n_merges) {
    t->nodes.clear();
    t->nodes.reserve(1024 + (size_t)n_merges * 4);
    t->nodes.emplace_back();
    t->max_len = 1;
    for (uint32_t b = 0; b < 256; b++) t->insert(&b, 1, b);  // lib.rs:155-157
    for (uint32_t i = 0; i < n_merges; i++) {                // lib.rs:159-161
        const uint64_t o = seq_off[i], e = seq_off[i + 1];
        if (e <= o) return (int)i + 1;
        for (uint64_t k = o; k < e; k++)
            if (seq[k] >= 256) return (int)i + 1;
        t->insert(seq + o, (size_t)(e - o), ids[i]);
        t->max_len = std::max<uint32_t>(t->max_len, (uint32_t)(e - o));
    }
    return 0;
}

void vocab_classes(const HostTrie &t, uint8_t cls[256], int *n_classes) {
    bool needs[256] = {false};
    for (size_t n = 0; n < t.nodes.size(); n++) {
        for (auto &kv : t.nodes[n].child) {
            if (kv.first >= 256) continue;
            const HostNode &c = t.nodes[kv.second];
            if (n != 0) needs[kv.first] = true;  // edge below depth 1
            else if (!c.child.empty() || c.token != (int64_t)kv.first) needs[kv.first] = true;
        }
    }
    std::memset(cls, 255, 256);
    for (int k = 0; k < kNumSymbols; k++) cls['a' + k] = (uint8_t)k;
    int nc = kNumSymbols;
    for (int b = 0; b < 256; b++) {
        if (!needs[b] || (b >= 'a' && b <= 'z')) continue;
        if (nc < 255) cls[b] = (uint8_t)nc;
        nc++;
    }
    *n_classes = nc;
}

namespace {
struct Row {
    int node;                                     // trie node of the state
    std::vector<std::pair<uint32_t, int>> slots;  // (code, target node); singles carry SM in the low field
    uint32_t base = 0;
};
}  // namespace

bool build_pairtab(const HostTrie &t, const uint8_t cls[256], int n_classes, PairTab *out) {
    PairTab &p = *out;
    p = PairTab();
    if (n_classes < 1) return false;
    p.NC = (uint32_t)n_classes;
    p.SM = p.NC;
    p.SE = p.NC + 1;
    p.W = 1;
    while ((1u << p.W) <= p.NC + 2) p.W++;  // NC + 2 < 2^W: the all-ones class is never used
    if (p.W > 6) return false;
    const uint32_t W = p.W;
    const uint32_t span = 1u << (2 * W);  // codes are < span

    // children of a node by class, classes ascending
    auto class_children = [&](int n, std::vector<std::pair<uint32_t, int>> &ch) {
        ch.clear();
        for (auto &kv : t.nodes[n].child)
            if (kv.first < 256 && cls[kv.first] < n_classes) ch.push_back({cls[kv.first], kv.second});
        std::sort(ch.begin(), ch.end());
    };

    // states = nodes at even depth, reached through class edges only
    std::vector<Row> rows;
    std::vector<int> state_of(t.nodes.size(), -1);
    std::vector<int> queue{0};
    state_of[0] = 0;
    rows.push_back(Row{0, {}, 0});
    std::vector<std::pair<uint32_t, int>> ch1, ch2;
    for (size_t head = 0; head < queue.size(); head++) {
        const int s = queue[head];
        const int ri = state_of[s];
        class_children(s, ch1);
        for (auto &c1 : ch1) {
            const HostNode &u = t.nodes[c1.second];
            if (u.token >= 0) {
                if (u.token > 0xFFFF) return false;
                rows[ri].slots.push_back({(c1.first << W) | p.SM, c1.second});
            }
            class_children(c1.second, ch2);
            for (auto &c2 : ch2) {
                const HostNode &w = t.nodes[c2.second];
                if (w.token > 0xFFFF) return false;
                rows[ri].slots.push_back({(c1.first << W) | c2.first, c2.second});
                if (state_of[c2.second] < 0) {
                    state_of[c2.second] = (int)rows.size();
                    rows.push_back(Row{c2.second, {}, 0});
                    queue.push_back(c2.second);
                }
            }
        }
    }
    p.n_states = (uint32_t)rows.size();

    // first-fit placement, large rows first; every non-empty row gets a base of its own
    std::vector<int> order;
    for (size_t i = 0; i < rows.size(); i++)
        if (!rows[i].slots.empty()) order.push_back((int)i);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return rows[x].slots.size() > rows[y].slots.size(); });
    size_t total = 0;
    for (int i : order) total += rows[i].slots.size();
    std::vector<uint8_t> used(total + 2 * span + 64, 0), base_used(total + 2 * span + 64, 0);
    size_t first_free = 0;  // every slot below it is taken
    size_t max_slot = 0;
    for (int i : order) {
        Row &r = rows[i];
        uint32_t min_code = r.slots[0].first;
        for (auto &s : r.slots) min_code = std::min(min_code, s.first);
        size_t b = first_free > min_code ? first_free - min_code : 0;
        for (;; b++) {
            if (b + span >= used.size()) { used.resize(used.size() * 2, 0); base_used.resize(used.size(), 0); }
            if (base_used[b]) continue;
            bool fits = true;
            for (auto &s : r.slots)
                if (used[b + s.first]) { fits = false; break; }
            if (fits) break;
        }
        r.base = (uint32_t)b;
        base_used[b] = 1;
        for (auto &s : r.slots) {
            used[b + s.first] = 1;
            max_slot = std::max(max_slot, b + s.first);
        }
        while (first_free < used.size() && used[first_free]) first_free++;
    }
    // the dead base: no row lives there, so no probe from it can match; probes stay in bounds
    size_t dead = max_slot + 1;
    while (dead < base_used.size() && base_used[dead]) dead++;
    const size_t n_ent = std::max(dead, max_slot + 1) + span;
    if (n_ent > 0xFFFF) return false;
    p.dead_base = (uint32_t)dead;
    p.root_base = rows[0].slots.empty() ? p.dead_base : rows[0].base;
    p.ent.assign(n_ent, 0xFFFFFFFFu);
    p.tok.assign(n_ent, 0);
    for (int i : order) {
        const Row &r = rows[i];
        for (auto &s : r.slots) {
            const uint32_t code = s.first;
            const HostNode &target = t.nodes[s.second];
            uint32_t e;
            if ((code & ((1u << W) - 1)) == p.SM) {
                e = (p.dead_base << 16) | (code << 2);
            } else {
                // the node between: child of the state by the high class
                int mid = -1;
                for (auto &kv : t.nodes[r.node].child)
                    if (kv.first < 256 && cls[kv.first] == (code >> W)) mid = kv.second;
                const uint32_t t1 = (mid >= 0 && t.nodes[mid].token >= 0) ? 1u : 0u;
                const uint32_t t2 = target.token >= 0 ? 1u : 0u;
                const Row &nr = rows[state_of[s.second]];
                const uint32_t nb = nr.slots.empty() ? p.dead_base : nr.base;
                e = (nb << 16) | (code << 2) | (t2 << 1) | t1;
            }
            p.ent[r.base + code] = e;
            p.tok[r.base + code] = target.token >= 0 ? (uint16_t)target.token : 0;
            p.n_slots_used++;
        }
    }
    return true;
}

}  // namespace ecgb

using namespace ecgb;

// Host-only view of the pair table of a merges list (inspection / tests): two-call sizing through
// *n_ent_out.  h_ent / h_tok may be NULL.  meta[8] = {root_base, dead_base, W, NC, SM, SE, n_states, slots used}.
extern "C" int ecgb_pairtab_host(const uint32_t *h_seq, const uint64_t *h_seq_off, const uint32_t *h_ids, uint32_t n_merges,
                                 uint32_t *h_ent, uint16_t *h_tok, uint32_t cap, uint32_t *n_ent_out, uint32_t meta[8],
                                 uint8_t h_cls_out[256]) {
    ECGB_REQUIRE(n_ent_out && meta, "NULL argument");
    ECGB_REQUIRE(n_merges == 0 || (h_seq && h_seq_off && h_ids), "NULL merge arrays");
    HostTrie t;
    const int bad = build_host_trie(&t, h_seq, h_seq_off, h_ids, n_merges);
    ECGB_REQUIRE(bad == 0, "merge %d is malformed", bad - 1);
    uint8_t cls[256];
    int n_classes = 0;
    vocab_classes(t, cls, &n_classes);
    PairTab p;
    if (!build_pairtab(t, cls, n_classes, &p)) return fail(ECGB_EUNSUPPORTED, "vocabulary does not fit the pair table");
    *n_ent_out = (uint32_t)p.ent.size();
    meta[0] = p.root_base; meta[1] = p.dead_base; meta[2] = p.W; meta[3] = p.NC;
    meta[4] = p.SM; meta[5] = p.SE; meta[6] = p.n_states; meta[7] = p.n_slots_used;
    if (h_cls_out) std::memcpy(h_cls_out, cls, 256);
    if (h_ent && h_tok) {
        if (cap < p.ent.size()) return fail(ECGB_ECAPACITY, "pair table has %zu slots", p.ent.size());
        std::memcpy(h_ent, p.ent.data(), p.ent.size() * 4);
        std::memcpy(h_tok, p.tok.data(), p.tok.size() * 2);
    }
    return ECGB_OK;
}
