// Native host side of the batched MCTS scheduler: bit-packed SimpleNetworkGrammar states,
// canonicalisation over component permutations, action generation, UCB selection, random
// rollouts, visit/reward back-propagation and exhaustion bookkeeping.
//
// It restates, on packed 64-bit states, exactly what the reference does on strings:
//   get_actions / do_action            reference circuitree/models.py:164-270
//   get_unique_state (min relabelling) reference circuitree/models.py:337-400
//   select_and_expand, ucb_score       reference circuitree/circuitree.py:21-57, :315-384
//   random rollout                     reference circuitree/circuitree.py:218-246
//   _backpropagate                     reference circuitree/circuitree.py:451-466
//   mark_as_exhausted                  reference circuitree/circuitree.py:386-420
//   grow_tree                          reference circuitree/circuitree.py:539-586
// including the consumption of numpy's PCG64 stream by Generator.shuffle / Generator.choice, so
// that for identical rewards the tree is identical to the reference's, node for node.
//
// Scope: all components present in the root (e.g. "ABC::"), at most 5 components and 3
// interaction types, no fixed components.  Anything else stays on the Python path.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/circuitree_b200.h"

namespace {

thread_local char t_err[512] = "";
int tfail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
    return code;
}

typedef unsigned __int128 u128;

struct NvtxRange {     // host-side phases of the scheduler on the profiler timeline
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// ---- numpy's PCG64 (XSL-RR 128/64) and the two bounded-integer algorithms the reference uses ----
struct Pcg64 {
    u128 state = 0, inc = 0;
    int has_uint32 = 0;
    uint32_t uinteger = 0;

    uint64_t next64()
    {
        const u128 mult = ((u128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
        state = state * mult + inc;
        const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
        const unsigned rot = (unsigned)(state >> 122);
        const uint64_t x = hi ^ lo;
        return (x >> rot) | (x << ((-rot) & 63));
    }
    uint32_t next32()
    {
        if (has_uint32) { has_uint32 = 0; return uinteger; }
        const uint64_t n = next64();
        has_uint32 = 1;
        uinteger = (uint32_t)(n >> 32);
        return (uint32_t)n;
    }
    // numpy random_interval(max): masked rejection, used by Generator.shuffle on a list
    uint64_t interval(uint64_t max)
    {
        if (max == 0) return 0;
        uint64_t mask = max, value;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
        if (max <= 0xffffffffull) { while ((value = (next32() & mask)) > max) {} }
        else { while ((value = (next64() & mask)) > max) {} }
        return value;
    }
    // numpy Generator.integers(0, n) for n <= 2^32 (Lemire, 32-bit buffered), used by Generator.choice
    uint64_t below(uint64_t n)
    {
        const uint64_t rng = n - 1;
        if (rng == 0) return 0;
        if (rng == 0xffffffffull) return next32();
        const uint32_t rng_excl = (uint32_t)rng + 1u;
        uint64_t m = (uint64_t)next32() * rng_excl;
        uint32_t leftover = (uint32_t)m;
        if (leftover < rng_excl) {
            const uint32_t threshold = (0xffffffffu - (uint32_t)rng) % rng_excl;
            while (leftover < threshold) {
                m = (uint64_t)next32() * rng_excl;
                leftover = (uint32_t)m;
            }
        }
        return m >> 32;
    }
    template <typename T>
    void shuffle(std::vector<T>& x)
    {
        for (size_t i = x.size(); i-- > 1;) {
            const size_t j = (size_t)interval(i);
            std::swap(x[i], x[j]);
        }
    }
};

constexpr uint64_t kTerminalBit = 1ull << 62;
constexpr int kMaxN = 5;

struct Node {
    uint64_t key;
    int64_t visits = 0;
    double reward = 0.0;
    bool exhausted = false;
    std::vector<int> out_edges, in_edges;   // insertion order, like networkx adjacency dicts
    int64_t sync_visits = 0;                // values at the last delta collection (multi-GPU merge)
    double sync_reward = 0.0;
};
struct Edge {
    int parent, child;
    int64_t visits = 0;
    double reward = 0.0;
    int64_t sync_visits = 0;
    double sync_reward = 0.0;
    int64_t unc_visits = 0;                 // part of (visits, reward) already un-counted from the parent
    double unc_reward = 0.0;                // because the child is flagged exhausted (root-parallel replicas)
};

// ---- DimersGrammar states (reference circuitree/models.py:524-939; semantics = this repo's Python mirror,
// circuitree_b200/models.py, which restates the documented intent where upstream raises) ----
// A state is the sorted list of its interactions; an interaction "m1 m2 type target" (monomers sorted) is a
// small integer code whose numeric order is the ASCII order of the four characters, so comparing code lists
// lexicographically is comparing the reference's interaction strings.  Lists are interned: a tree key is the
// list's id (+ the terminal bit), which keeps the tree code of the pairwise grammar unchanged.
struct DimerSpace {
    int S = 0, C = 0, T = 0;                 // species (components + regulators), components, interaction types
    char letter[kMaxN];                      // species rank (ASCII order) -> letter
    bool is_comp[kMaxN];
    int comp_pos[kMaxN], reg_pos[kMaxN];     // species rank -> index in the sorted components / regulators string
    int decl_comp[kMaxN];                    // component declaration index -> species rank
    char type_letter[3];                     // type rank (ASCII order) -> letter
    int decl_type[3];                        // declaration index -> type rank
    int max_interactions = 0, max_per_promoter = 2;
    std::vector<std::pair<int, int>> dimers; // valid dimers (m1 <= m2 as species ranks) in ASCII order of the two letters
    std::vector<std::vector<uint16_t>> remap;    // relabelling (component perm x regulator perm) -> code -> code
    std::vector<std::vector<uint16_t>> lists;    // interned interaction lists (sorted codes)
    std::unordered_map<std::string, uint32_t> ids;
    std::unordered_map<uint32_t, uint32_t> canon_memo;
    std::string comps_sorted, regs_sorted;

    int code_of(int m1, int m2, int type, int target) const { return ((m1 * S + m2) * T + type) * S + target; }
    void decode(int code, int& m1, int& m2, int& type, int& target) const
    {
        target = code % S; code /= S;
        type = code % T; code /= T;
        m2 = code % S; m1 = code / S;
    }
    uint32_t intern(const std::vector<uint16_t>& v)
    {
        std::string k(reinterpret_cast<const char*>(v.data()), v.size() * sizeof(uint16_t));
        auto it = ids.find(k);
        if (it != ids.end()) return it->second;
        const uint32_t id = (uint32_t)lists.size();
        lists.push_back(v);
        ids.emplace(std::move(k), id);
        return id;
    }
    uint32_t canonical(uint32_t id)
    {
        auto it = canon_memo.find(id);
        if (it != canon_memo.end()) return it->second;
        const std::vector<uint16_t> src = lists[id];
        std::vector<uint16_t> best, cur(src.size());
        bool first = true;
        for (const auto& map : remap) {
            for (size_t x = 0; x < src.size(); ++x) cur[x] = map[src[x]];
            std::sort(cur.begin(), cur.end());
            if (first || cur < best) { best = cur; first = false; }
        }
        const uint32_t out = intern(best);
        canon_memo.emplace(id, out);
        return out;
    }
    // action ids: 0 = *terminate*, 1 + ((dimer * C + declared component) * T + declared type)
    void actions(uint64_t key, std::vector<int>& out) const
    {
        out.clear();
        if (key & kTerminalBit) return;
        const std::vector<uint16_t>& v = lists[(size_t)(key & 0xffffffffull)];
        // triplets (dimer, target) in use, letters in use, load per promoter
        bool in_use[kMaxN] = {false, false, false, false, false};
        int load[kMaxN] = {0, 0, 0, 0, 0};
        std::vector<int> trip;
        for (uint16_t c : v) {
            int m1, m2, ty, tg;
            decode(c, m1, m2, ty, tg);
            in_use[m1] = in_use[m2] = in_use[tg] = true;
            load[tg] += 1;
            const int t3 = (m1 * S + m2) * S + tg;
            if (std::find(trip.begin(), trip.end(), t3) == trip.end()) trip.push_back(t3);
        }
        if ((int)trip.size() >= max_interactions) { out.push_back(0); return; }
        if (!trip.empty()) out.push_back(0);
        for (size_t d = 0; d < dimers.size(); ++d)
            for (int dc = 0; dc < C; ++dc) {
                const int m1 = dimers[d].first, m2 = dimers[d].second, tg = decl_comp[dc];
                if (!trip.empty()) {
                    if (load[tg] >= max_per_promoter) continue;
                    if (std::find(trip.begin(), trip.end(), (m1 * S + m2) * S + tg) != trip.end()) continue;
                    if (!in_use[m1] && !in_use[m2] && !in_use[tg]) continue;
                }
                for (int t = 0; t < T; ++t) out.push_back(1 + ((int)d * C + dc) * T + t);
            }
    }
    uint64_t apply(uint64_t key, int action)
    {
        if (action == 0) return key | kTerminalBit;
        const int a = action - 1, t = a % T, dc = (a / T) % C, d = a / (T * C);
        std::vector<uint16_t> v = lists[(size_t)(key & 0xffffffffull)];
        v.push_back((uint16_t)code_of(dimers[(size_t)d].first, dimers[(size_t)d].second, decl_type[t], decl_comp[dc]));
        std::sort(v.begin(), v.end());
        return (key & kTerminalBit) | intern(v);
    }
    uint64_t step(uint64_t key, int action)
    {
        const uint64_t k = apply(key, action);
        return (k & kTerminalBit) | canonical((uint32_t)(k & 0xffffffffull));
    }
    std::string to_string(uint64_t key) const
    {
        std::string s;
        if (key & kTerminalBit) s += '*';
        s += comps_sorted; s += '+'; s += regs_sorted; s += "::";
        const std::vector<uint16_t>& v = lists[(size_t)(key & 0xffffffffull)];
        for (size_t x = 0; x < v.size(); ++x) {
            int m1, m2, ty, tg;
            decode(v[x], m1, m2, ty, tg);
            if (x) s += '_';
            s += letter[m1]; s += letter[m2]; s += type_letter[ty]; s += letter[tg];
        }
        return s;
    }
    int rank_of(char c) const
    {
        for (int i = 0; i < S; ++i) if (letter[i] == c) return i;
        return -1;
    }
    bool from_string(const char* str, uint64_t* out)
    {
        const char* p = str;
        uint64_t flags = 0;
        if (*p == '*') { flags = kTerminalBit; ++p; }
        const std::string head = comps_sorted + "+" + regs_sorted + "::";
        // components and regulators may come in any order, but all must be present
        const char* sep = strstr(p, "::");
        if (!sep) return false;
        std::string given(p, sep);
        const size_t plus = given.find('+');
        if (plus == std::string::npos) return false;
        std::string gc = given.substr(0, plus), gr = given.substr(plus + 1);
        std::sort(gc.begin(), gc.end()); std::sort(gr.begin(), gr.end());
        if (gc != comps_sorted || gr != regs_sorted) return false;
        p = sep + 2;
        std::vector<uint16_t> v;
        while (*p) {
            if (strlen(p) < 4) return false;
            int m1 = rank_of(p[0]), m2 = rank_of(p[1]), tg = rank_of(p[3]), ty = -1;
            for (int t = 0; t < T; ++t) if (type_letter[t] == p[2]) ty = t;
            if (m1 < 0 || m2 < 0 || tg < 0 || ty < 0 || !is_comp[tg]) return false;
            if (m1 > m2) std::swap(m1, m2);
            v.push_back((uint16_t)code_of(m1, m2, ty, tg));
            p += 4;
            if (*p == '_') ++p; else if (*p) return false;
        }
        std::sort(v.begin(), v.end());
        *out = flags | intern(v);
        return true;
    }
    // GPU topology handle (ct_compile_topology, k = 3): rows (monomer1, monomer2, target) index components + regulators
    bool to_handle(uint64_t key, uint64_t* handle) const
    {
        const std::vector<uint16_t>& v = lists[(size_t)(key & 0xffffffffull)];
        std::vector<int64_t> act, inh;
        for (uint16_t c : v) {
            int m1, m2, ty, tg;
            decode(c, m1, m2, ty, tg);
            auto pos = [&](int r) { return (int64_t)(is_comp[r] ? comp_pos[r] : C + reg_pos[r]); };
            int64_t a = pos(m1), b = pos(m2);
            if (a > b) std::swap(a, b);
            std::vector<int64_t>& dst = type_letter[ty] == 'a' ? act : inh;
            if (type_letter[ty] != 'a' && type_letter[ty] != 'i') return false;
            dst.push_back(a); dst.push_back(b); dst.push_back(pos(tg));
        }
        return ct_compile_topology(S, act.data(), (int32_t)(act.size() / 3), inh.data(), (int32_t)(inh.size() / 3), 3, handle) == CT_OK;
    }
    int depth_of(uint64_t key) const { return (int)lists[(size_t)(key & 0xffffffffull)].size() + ((key & kTerminalBit) ? 1 : 0); }

    // Keys are process-local ids of interned lists, so replicas exchange the lists themselves: a state travels as
    // wire_words() 64-bit words -- word 0: terminal flag (bit 0), list length (bits 1..7), codes 0..5 (9 bits each from
    // bit 8); every further word: seven more codes from bit 0.  Every replica of a search builds the same code table
    // (same components, regulators and types), and node keys are canonical lists, so the codes mean the same everywhere.
    int max_list_len() const
    {
        const int by_promoter = max_per_promoter * C;
        return max_interactions < by_promoter ? max_interactions : by_promoter;
    }
    int wire_words() const
    {
        const int cap = max_list_len();
        return 1 + (cap > 6 ? (cap - 6 + 6) / 7 : 0);
    }
    bool to_wire(uint64_t key, int64_t* w) const
    {
        const std::vector<uint16_t>& v = lists[(size_t)(key & 0xffffffffull)];
        const int W = wire_words();
        if ((int)v.size() > 6 + 7 * (W - 1) || v.size() > 127) return false;
        for (int x = 0; x < W; ++x) w[x] = 0;
        uint64_t w0 = ((key & kTerminalBit) ? 1ull : 0ull) | ((uint64_t)v.size() << 1);
        for (size_t x = 0; x < v.size(); ++x) {
            if (v[x] >= 512) return false;
            if (x < 6) w0 |= (uint64_t)v[x] << (8 + 9 * x);
            else w[1 + (x - 6) / 7] |= (int64_t)((uint64_t)v[x] << (9 * ((x - 6) % 7)));
        }
        w[0] = (int64_t)w0;
        return true;
    }
    bool from_wire(const int64_t* w, uint64_t* key)
    {
        const uint64_t w0 = (uint64_t)w[0];
        const size_t n = (size_t)((w0 >> 1) & 127u);
        if ((int)n > 6 + 7 * (wire_words() - 1)) return false;
        std::vector<uint16_t> v(n);
        const int n_codes = S * S * T * S;
        for (size_t x = 0; x < n; ++x) {
            v[x] = (uint16_t)(x < 6 ? (w0 >> (8 + 9 * x)) & 511u : ((uint64_t)w[1 + (x - 6) / 7] >> (9 * ((x - 6) % 7))) & 511u);
            if (v[x] >= n_codes || (x && v[x] <= v[x - 1])) return false;      // a sorted list of distinct valid codes
        }
        *key = ((w0 & 1u) ? kTerminalBit : 0ull) | intern(v);
        return true;
    }
};

}  // namespace

struct ct_tree {
    DimerSpace* dim = nullptr;         // non-null: a DimersGrammar tree (keys are interned list ids)
    ~ct_tree() { delete dim; }
    int n = 0, n_types = 0, n_cells = 0, max_interactions = 0;
    char comp_letter[kMaxN];       // sorted index -> letter
    int decl_comp[kMaxN];          // declaration index -> sorted index
    char type_letter[3];           // type code (rank by letter) -> letter
    int decl_type[3];              // declaration index -> type code
    std::vector<std::vector<uint8_t>> perm_cell;   // perm -> cell -> cell
    uint64_t all_absent = 0;
    double c_ucb = std::sqrt(2.0);
    bool track_exhaustion = false;
    double n_exhausted = 0.0;
    Pcg64 rng;
    std::vector<double> log_table;
    uint64_t root_key = 0;
    std::vector<Node> nodes;
    std::vector<Edge> edges;
    std::unordered_map<uint64_t, int> index;
    std::unordered_map<uint64_t, uint64_t> canon_memo;
    std::vector<std::vector<int>> pending;   // paths (node ids) awaiting rewards, FIFO
    size_t pending_head = 0;
    bool exhausted_all = false;
    std::vector<uint64_t> newly_exhausted;   // keys flagged by this replica since the last delta collection
    std::vector<int> exhausted_nodes;        // every flagged node, in flagging order

    int shift(int cell) const { return 2 * (n_cells - 1 - cell); }
    uint32_t cell(uint64_t key, int c) const { return (uint32_t)(key >> shift(c)) & 3u; }

    uint64_t canonical(uint64_t key)
    {
        if (dim) return (key & kTerminalBit) | dim->canonical((uint32_t)(key & 0xffffffffull));
        const uint64_t flags = key & kTerminalBit;
        const uint64_t body = key & ~kTerminalBit;
        auto it = canon_memo.find(body);
        if (it != canon_memo.end()) return it->second | flags;
        int cells[kMaxN * kMaxN], codes[kMaxN * kMaxN], ne = 0;
        for (int c = 0; c < n_cells; ++c) {
            const uint32_t v = cell(body, c);
            if (v != 3u) { cells[ne] = c; codes[ne] = (int)v; ++ne; }
        }
        uint64_t best = ~0ull;
        for (const auto& pm : perm_cell) {
            uint64_t k = all_absent;
            for (int e = 0; e < ne; ++e) k ^= (uint64_t)(3u ^ (uint32_t)codes[e]) << shift(pm[cells[e]]);
            if (k < best) best = k;
        }
        if (canon_memo.size() > (1u << 22)) canon_memo.clear();
        canon_memo.emplace(body, best);
        return best | flags;
    }

    // action ids: 0 = *terminate*, 1 + (decl_pair * n_types + decl_type) = add interaction
    void actions(uint64_t key, std::vector<int>& out) const
    {
        if (dim) { dim->actions(key, out); return; }
        out.clear();
        if (key & kTerminalBit) return;
        int n_ixn = 0;
        bool touched[kMaxN] = {false, false, false, false, false};
        for (int c = 0; c < n_cells; ++c)
            if (cell(key, c) != 3u) { ++n_ixn; touched[c / n] = true; touched[c % n] = true; }
        if (n_ixn >= max_interactions) { out.push_back(0); return; }
        if (n_ixn > 0) out.push_back(0);
        for (int di = 0; di < n; ++di)
            for (int dj = 0; dj < n; ++dj) {
                const int si = decl_comp[di], sj = decl_comp[dj];
                if (n_ixn > 0) {
                    if (cell(key, si * n + sj) != 3u) continue;
                    if (!touched[si] && !touched[sj]) continue;
                }
                for (int t = 0; t < n_types; ++t) out.push_back(1 + (di * n + dj) * n_types + t);
            }
    }
    uint64_t apply(uint64_t key, int action) const
    {
        if (action == 0) return key | kTerminalBit;
        const int a = action - 1, t = a % n_types, pair = a / n_types;
        const int si = decl_comp[pair / n], sj = decl_comp[pair % n];
        const int sh = shift(si * n + sj);
        return (key & ~(3ull << sh)) | ((uint64_t)decl_type[t] << sh);
    }
    uint64_t step(uint64_t key, int action) { return dim ? dim->step(key, action) : canonical(apply(key, action)); }

    int node_of(uint64_t key) const
    {
        auto it = index.find(key);
        return it == index.end() ? -1 : it->second;
    }
    int add_node(uint64_t key)
    {
        const int id = (int)nodes.size();
        nodes.emplace_back();
        nodes.back().key = key;
        index.emplace(key, id);
        return id;
    }
    int edge_of(int parent, int child) const
    {
        for (int e : nodes[parent].out_edges)
            if (edges[e].child == child) return e;
        return -1;
    }
    int add_edge(int parent, int child)
    {
        const int id = (int)edges.size();
        edges.emplace_back();
        edges.back().parent = parent;
        edges.back().child = child;
        nodes[parent].out_edges.push_back(id);
        nodes[child].in_edges.push_back(id);
        return id;
    }
    double log_of(int64_t v) const
    {
        if (v >= 0 && (size_t)v < log_table.size()) return log_table[(size_t)v];
        return std::log((double)v);
    }

    // returns false when the whole tree is exhausted (the reference raises ExhaustionError)
    // Root-parallel replicas (ct_tree_pack_deltas / ct_tree_merge_packed): the un-counting below is a
    // consequence of the exhaustion EVENT, which every replica replays for itself; it must not travel
    // as a statistic increment as well, so the sync baseline moves with it.  `record` is false only
    // for the node another replica's event names (nodes flagged by the recursion are always recorded:
    // receivers skip what they have flagged already).
    bool mark_exhausted(int node, bool record = true)
    {
        nodes[node].exhausted = true;
        exhausted_nodes.push_back(node);
        if (record) newly_exhausted.push_back(nodes[node].key);
        if (nodes[node].key == root_key) { exhausted_all = true; return false; }
        for (int e : nodes[node].in_edges) {
            Node& p = nodes[edges[e].parent];
            p.visits -= edges[e].visits;
            p.reward -= edges[e].reward;
            p.sync_visits -= edges[e].visits;
            p.sync_reward -= edges[e].reward;
            edges[e].unc_visits = edges[e].visits;
            edges[e].unc_reward = edges[e].reward;
        }
        for (size_t x = 0; x < nodes[node].in_edges.size(); ++x) {
            const int parent = edges[nodes[node].in_edges[x]].parent;
            bool all = true;
            for (int e : nodes[parent].out_edges)
                if (!nodes[edges[e].child].exhausted) { all = false; break; }
            if (all && !nodes[parent].exhausted && !mark_exhausted(parent)) return false;
        }
        return true;
    }

    // Root-parallel replicas only (never called by a single tree, which follows the reference to the
    // letter): whatever reached an edge into a flagged node after the node was flagged -- the
    // flagging traversal's own visit and reward, rewards of leaves that were in flight, other
    // replicas' increments merged later -- is un-counted from the parent too.  Every replica then
    // holds, for every parent, the sum of all increments minus everything that came through flagged
    // children, whichever replica did the flagging and whenever the increments arrived.
    void settle_exhausted()
    {
        for (int node : exhausted_nodes)
            for (int e : nodes[node].in_edges) {
                Edge& ed = edges[e];
                if (ed.visits == ed.unc_visits && ed.reward == ed.unc_reward) continue;
                Node& p = nodes[ed.parent];
                const int64_t dv = ed.visits - ed.unc_visits;
                const double dr = ed.reward - ed.unc_reward;
                p.visits -= dv; p.sync_visits -= dv;
                p.reward -= dr; p.sync_reward -= dr;
                ed.unc_visits = ed.visits; ed.unc_reward = ed.reward;
            }
    }

    // number of interactions of a state (+1 for a terminal state): its depth below the root
    int depth_of(uint64_t key) const
    {
        if (dim) return dim->depth_of(key);
        int d = (key & kTerminalBit) ? 1 : 0;
        for (int c = 0; c < n_cells; ++c) d += cell(key, c) != 3u;
        return d;
    }

    // 0 ok, 1 tree exhausted, 2 dead end (every child exhausted or NaN scores: the reference crashes here)
    int select_and_expand(std::vector<int>& path)
    {
        std::vector<int> acts;
        int node = node_of(root_key);
        path.assign(1, node);
        actions(nodes[node].key, acts);
        const double inf = std::numeric_limits<double>::infinity();
        while (!acts.empty()) {
            rng.shuffle(acts);
            int best = -1;
            double best_score = -inf;
            for (int a : acts) {
                const uint64_t ck = step(nodes[node].key, a);
                int child = node_of(ck);
                if (track_exhaustion && child >= 0 && nodes[child].exhausted) continue;
                const int e = child >= 0 ? edge_of(node, child) : -1;
                double score = inf;
                if (e >= 0 && edges[e].visits != 0)
                    score = edges[e].reward / (double)edges[e].visits +
                            c_ucb * std::sqrt(log_of(nodes[node].visits) / (double)edges[e].visits);
                if (score == inf) {
                    if (child < 0) child = add_node(ck);
                    if (e < 0) add_edge(node, child);
                    path.push_back(child);
                    return 0;
                }
                if (score > best_score) { best = child; best_score = score; }
            }
            if (best < 0) return 2;
            node = best;
            path.push_back(node);
            actions(nodes[node].key, acts);
        }
        if (track_exhaustion && (double)nodes[node].visits >= n_exhausted - 1.0)
            if (!mark_exhausted(node)) return 1;
        return 0;
    }

    uint64_t rollout(uint64_t key)
    {
        std::vector<int> acts;
        while (!(key & kTerminalBit)) {
            actions(key, acts);
            key = step(key, acts[(size_t)rng.below(acts.size())]);
        }
        return key;
    }

    template <typename F>
    void along(const std::vector<int>& path, F f)
    {
        int child = path.back();
        f(nodes[child].visits, nodes[child].reward);
        for (size_t x = path.size() - 1; x-- > 0;) {
            const int parent = path[x];
            Edge& e = edges[edge_of(parent, child)];
            f(e.visits, e.reward);
            f(nodes[parent].visits, nodes[parent].reward);
            child = parent;
        }
    }

    std::string to_string(uint64_t key) const
    {
        if (dim) return dim->to_string(key);
        std::string s;
        if (key & kTerminalBit) s += '*';
        for (int i = 0; i < n; ++i) s += comp_letter[i];
        s += "::";
        bool first = true;
        for (int c = 0; c < n_cells; ++c) {
            const uint32_t v = cell(key, c);
            if (v == 3u) continue;
            if (!first) s += '_';
            first = false;
            s += comp_letter[c / n];
            s += comp_letter[c % n];
            s += type_letter[v];
        }
        return s;
    }
    bool from_string(const char* str, uint64_t* out)
    {
        if (dim) return dim->from_string(str, out);
        const char* p = str;
        uint64_t key = all_absent;
        if (*p == '*') { key |= kTerminalBit; ++p; }
        const char* sep = strstr(p, "::");
        if (!sep) return false;
        // every component must be present (any order)
        if ((int)(sep - p) != n) return false;
        for (int i = 0; i < n; ++i)
            if (!memchr(p, comp_letter[i], (size_t)n)) return false;
        p = sep + 2;
        while (*p) {
            if (strlen(p) < 3) return false;
            int si = -1, sj = -1, tc = -1;
            for (int i = 0; i < n; ++i) { if (comp_letter[i] == p[0]) si = i; if (comp_letter[i] == p[1]) sj = i; }
            for (int t = 0; t < n_types; ++t) if (type_letter[t] == p[2]) tc = t;
            if (si < 0 || sj < 0 || tc < 0) return false;
            const int sh = shift(si * n + sj);
            key = (key & ~(3ull << sh)) | ((uint64_t)tc << sh);
            p += 3;
            if (*p == '_') ++p; else if (*p) return false;
        }
        *out = key;
        return true;
    }
    // GPU topology handle (ct_compile_topology layout): type letter 'a' -> 1, 'i' -> 2
    bool to_handle(uint64_t key, uint64_t* handle) const
    {
        if (dim) return dim->to_handle(key, handle);
        uint64_t h = (uint64_t)n << 56;
        for (int c = 0; c < n_cells; ++c) {
            const uint32_t v = cell(key, c);
            if (v == 3u) continue;
            const char t = type_letter[v];
            const uint64_t code = t == 'a' ? 1u : (t == 'i' ? 2u : 0u);
            if (!code) return false;
            h |= code << (2 * ((c / n) * 5 + (c % n)));
        }
        *handle = h;
        return true;
    }
};

extern "C" {

const char* ct_tree_last_error(void) { return t_err; }

int ct_tree_create(const char* components, const char* types, int32_t max_interactions, const char* root,
                   double exploration_constant, double n_exhausted, ct_tree** out)
{
    if (!components || !types || !root || !out) return tfail(CT_ERR_INVALID, "NULL argument");
    const int n = (int)strlen(components), nt = (int)strlen(types);
    if (n < 1 || n > kMaxN) return tfail(CT_ERR_UNSUPPORTED, "native tree supports 1..5 components, got %d", n);
    if (nt < 1 || nt > 3) return tfail(CT_ERR_UNSUPPORTED, "native tree supports 1..3 interaction types, got %d", nt);
    ct_tree* t = new ct_tree();
    t->n = n; t->n_types = nt; t->n_cells = n * n;
    t->max_interactions = max_interactions < 0 ? n * n : max_interactions;
    std::string sc(components), st(types);
    std::sort(sc.begin(), sc.end());
    std::sort(st.begin(), st.end());
    if (std::adjacent_find(sc.begin(), sc.end()) != sc.end() || std::adjacent_find(st.begin(), st.end()) != st.end()) {
        delete t;
        return tfail(CT_ERR_INVALID, "component / interaction codes must be unique");
    }
    for (int i = 0; i < n; ++i) { t->comp_letter[i] = sc[i]; t->decl_comp[i] = (int)sc.find(components[i]); }
    for (int i = 0; i < nt; ++i) { t->type_letter[i] = st[i]; t->decl_type[i] = (int)st.find(types[i]); }
    t->all_absent = 0;
    for (int c = 0; c < t->n_cells; ++c) t->all_absent |= 3ull << t->shift(c);
    std::vector<int> p(n);
    for (int i = 0; i < n; ++i) p[i] = i;
    do {
        std::vector<uint8_t> m((size_t)t->n_cells);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) m[i * n + j] = (uint8_t)(p[i] * n + p[j]);
        t->perm_cell.push_back(m);
    } while (std::next_permutation(p.begin(), p.end()));
    t->c_ucb = exploration_constant;
    t->track_exhaustion = !std::isnan(n_exhausted);
    t->n_exhausted = n_exhausted;
    uint64_t rk;
    if (!t->from_string(root, &rk)) {
        delete t;
        return tfail(CT_ERR_UNSUPPORTED, "root '%s' must contain every component exactly once", root);
    }
    if (t->to_string(rk) != root) {
        delete t;
        return tfail(CT_ERR_UNSUPPORTED, "root '%s' is not in sorted form (expected '%s')", root, t->to_string(rk).c_str());
    }
    t->root_key = rk;
    t->add_node(rk);
    *out = t;
    return CT_OK;
}

/* A tree over DimersGrammar states (reference circuitree/models.py:524-939; semantics of this repo's Python
 * mirror).  components / regulators / types: one character each, in declaration order; the root holds every
 * component and regulator in sorted form ("ABC+DE::").  At most 5 TFs in total. */
int ct_tree_create_dimers(const char* components, const char* regulators, const char* types, int32_t max_interactions,
                          int32_t max_per_promoter, const char* root, double exploration_constant, double n_exhausted,
                          ct_tree** out)
{
    if (!components || !regulators || !types || !root || !out) return tfail(CT_ERR_INVALID, "NULL argument");
    const int C = (int)strlen(components), R = (int)strlen(regulators), T = (int)strlen(types), S = C + R;
    if (C < 1 || S > kMaxN) return tfail(CT_ERR_UNSUPPORTED, "native dimer tree supports 1..5 TFs in total, got %d + %d", C, R);
    if (T < 1 || T > 3) return tfail(CT_ERR_UNSUPPORTED, "native tree supports 1..3 interaction types, got %d", T);
    std::string all = std::string(components) + regulators, st(types);
    std::sort(all.begin(), all.end());
    std::sort(st.begin(), st.end());
    if (std::adjacent_find(all.begin(), all.end()) != all.end() || std::adjacent_find(st.begin(), st.end()) != st.end())
        return tfail(CT_ERR_INVALID, "component / regulator / interaction codes must be unique");
    ct_tree* t = new ct_tree();
    DimerSpace* D = t->dim = new DimerSpace();
    D->S = S; D->C = C; D->T = T;
    D->comps_sorted = components; D->regs_sorted = regulators;
    std::sort(D->comps_sorted.begin(), D->comps_sorted.end());
    std::sort(D->regs_sorted.begin(), D->regs_sorted.end());
    for (int i = 0; i < S; ++i) {
        D->letter[i] = all[(size_t)i];
        const size_t cp = D->comps_sorted.find(all[(size_t)i]);
        D->is_comp[i] = cp != std::string::npos;
        D->comp_pos[i] = D->is_comp[i] ? (int)cp : -1;
        D->reg_pos[i] = D->is_comp[i] ? -1 : (int)D->regs_sorted.find(all[(size_t)i]);
    }
    for (int i = 0; i < C; ++i) D->decl_comp[i] = (int)all.find(components[i]);
    for (int i = 0; i < T; ++i) { D->type_letter[i] = st[(size_t)i]; D->decl_type[i] = (int)st.find(types[i]); }
    D->max_interactions = max_interactions < 0 ? C * S * S : max_interactions;
    D->max_per_promoter = max_per_promoter;
    // dimers in ASCII order: both monomers components, or one component and one regulator (reference models.py:596-612)
    for (int a = 0; a < S; ++a)
        for (int b = a; b < S; ++b)
            if (D->is_comp[a] || D->is_comp[b]) D->dimers.emplace_back(a, b);
    // relabellings: every permutation of the components combined with every permutation of the regulators
    std::vector<int> cs, rs;
    for (int i = 0; i < S; ++i) (D->is_comp[i] ? cs : rs).push_back(i);
    std::vector<int> cperm = cs;
    const int n_codes = S * S * T * S;
    do {
        std::vector<int> rperm = rs;
        do {
            int to[kMaxN];
            for (size_t x = 0; x < cs.size(); ++x) to[cs[x]] = cperm[x];
            for (size_t x = 0; x < rs.size(); ++x) to[rs[x]] = rperm[x];
            std::vector<uint16_t> map((size_t)n_codes, 0);
            for (int code = 0; code < n_codes; ++code) {
                int m1, m2, ty, tg;
                D->decode(code, m1, m2, ty, tg);
                int a = to[m1], b = to[m2];
                if (a > b) std::swap(a, b);
                map[(size_t)code] = (uint16_t)D->code_of(a, b, ty, to[tg]);
            }
            D->remap.push_back(std::move(map));
        } while (std::next_permutation(rperm.begin(), rperm.end()));
    } while (std::next_permutation(cperm.begin(), cperm.end()));
    t->n = S; t->n_types = T;
    t->c_ucb = exploration_constant;
    t->track_exhaustion = !std::isnan(n_exhausted);
    t->n_exhausted = n_exhausted;
    uint64_t rk;
    if (!D->from_string(root, &rk)) {
        delete t;
        return tfail(CT_ERR_UNSUPPORTED, "root '%s' must contain every component and regulator exactly once", root);
    }
    if (D->to_string(rk) != root) {
        const std::string want = D->to_string(rk);
        delete t;
        return tfail(CT_ERR_UNSUPPORTED, "root '%s' is not in sorted form (expected '%s')", root, want.c_str());
    }
    t->root_key = rk;
    t->add_node(rk);
    *out = t;
    return CT_OK;
}

int ct_tree_destroy(ct_tree* t) { delete t; return CT_OK; }

int ct_tree_set_rng(ct_tree* t, const uint64_t state[2], const uint64_t inc[2], int32_t has_uint32, uint32_t uinteger)
{
    if (!t || !state || !inc) return tfail(CT_ERR_INVALID, "NULL argument");
    t->rng.state = ((u128)state[1] << 64) | state[0];
    t->rng.inc = ((u128)inc[1] << 64) | inc[0];
    t->rng.has_uint32 = has_uint32;
    t->rng.uinteger = uinteger;
    return CT_OK;
}

int ct_tree_get_rng(ct_tree* t, uint64_t state[2], uint64_t inc[2], int32_t* has_uint32, uint32_t* uinteger)
{
    if (!t || !state || !inc || !has_uint32 || !uinteger) return tfail(CT_ERR_INVALID, "NULL argument");
    state[0] = (uint64_t)t->rng.state; state[1] = (uint64_t)(t->rng.state >> 64);
    inc[0] = (uint64_t)t->rng.inc; inc[1] = (uint64_t)(t->rng.inc >> 64);
    *has_uint32 = t->rng.has_uint32;
    *uinteger = t->rng.uinteger;
    return CT_OK;
}

int ct_tree_set_log_table(ct_tree* t, const double* table, int64_t n)
{
    if (!t || (n && !table) || n < 0) return tfail(CT_ERR_INVALID, "bad log table");
    t->log_table.assign(table, table + n);
    return CT_OK;
}

/* Test hook: draw from the RNG like Generator.shuffle (kind 0: returns the permutation of 0..n-1)
 * or Generator.integers(0, n) (kind 1: out[0]). */
int ct_tree_rng_draw(ct_tree* t, int32_t kind, int64_t n, int64_t* out)
{
    if (!t || !out || n < 1) return tfail(CT_ERR_INVALID, "bad arguments");
    if (kind == 0) {
        std::vector<int64_t> v((size_t)n);
        for (int64_t i = 0; i < n; ++i) v[(size_t)i] = i;
        t->rng.shuffle(v);
        memcpy(out, v.data(), sizeof(int64_t) * (size_t)n);
    } else {
        out[0] = (int64_t)t->rng.below((uint64_t)n);
    }
    return CT_OK;
}

/* Select k leaves under virtual loss: select_and_expand, random rollout, visit back-propagation
 * (the order of reference circuitree.py:529-533).  sim_keys / sim_handles (nullable) receive the
 * packed terminal state and the GPU topology handle of each simulated state.  Returns the number
 * selected in *n_done (fewer than k only when the tree is exhausted). */
int ct_tree_select(ct_tree* t, int32_t k, uint64_t* sim_keys, uint64_t* sim_handles, int32_t* n_done)
{
    NvtxRange nvtx_range("ct_tree_select");
    if (!t || k < 0 || !n_done) return tfail(CT_ERR_INVALID, "bad arguments");
    *n_done = 0;
    std::vector<int> path;
    for (int x = 0; x < k; ++x) {
        if (t->exhausted_all) return tfail(100, "ExhaustionError: every node in the tree has been visited to exhaustion");
        const int rc = t->select_and_expand(path);
        if (rc == 1) return tfail(100, "ExhaustionError: every node in the tree has been visited to exhaustion");
        if (rc == 2) return tfail(CT_ERR_INVALID, "selection reached a node without selectable children");
        const uint64_t sim = t->rollout(t->nodes[path.back()].key);
        t->along(path, [](int64_t& v, double&) { v += 1; });
        if (sim_keys) sim_keys[x] = sim;
        if (sim_handles && !t->to_handle(sim, &sim_handles[x]))
            return tfail(CT_ERR_UNSUPPORTED, "state has interaction types other than a/i");
        t->pending.push_back(path);
        *n_done = x + 1;
    }
    return CT_OK;
}

/* Back-propagate rewards for the k oldest pending selections (FIFO). */
int ct_tree_backprop(ct_tree* t, int32_t k, const double* rewards)
{
    NvtxRange nvtx_range("ct_tree_backprop");
    if (!t || k < 0 || (k && !rewards)) return tfail(CT_ERR_INVALID, "bad arguments");
    if ((size_t)k > t->pending.size() - t->pending_head) return tfail(CT_ERR_INVALID, "more rewards than pending selections");
    for (int x = 0; x < k; ++x) {
        const double r = rewards[x];
        t->along(t->pending[t->pending_head + x], [r](int64_t&, double& w) { w += r; });
    }
    t->pending_head += (size_t)k;
    if (t->pending_head == t->pending.size()) { t->pending.clear(); t->pending_head = 0; }
    return CT_OK;
}

/* Paths of the pending selections: lengths[k] and node ids concatenated (nullable ids to size). */
int ct_tree_pending(ct_tree* t, int32_t* n_pending, int32_t* lengths, int32_t* ids)
{
    if (!t || !n_pending) return tfail(CT_ERR_INVALID, "bad arguments");
    const size_t np = t->pending.size() - t->pending_head;
    *n_pending = (int32_t)np;
    size_t off = 0;
    for (size_t x = 0; x < np; ++x) {
        const auto& p = t->pending[t->pending_head + x];
        if (lengths) lengths[x] = (int32_t)p.size();
        if (ids) for (size_t y = 0; y < p.size(); ++y) ids[off + y] = p[y];
        off += p.size();
    }
    return CT_OK;
}

int ct_tree_size(ct_tree* t, int64_t* n_nodes, int64_t* n_edges)
{
    if (!t) return tfail(CT_ERR_INVALID, "NULL tree");
    if (n_nodes) *n_nodes = (int64_t)t->nodes.size();
    if (n_edges) *n_edges = (int64_t)t->edges.size();
    return CT_OK;
}

int ct_tree_export(ct_tree* t, uint64_t* node_keys, int64_t* node_visits, double* node_reward, uint8_t* node_exhausted,
                   int32_t* edge_parent, int32_t* edge_child, int64_t* edge_visits, double* edge_reward)
{
    if (!t) return tfail(CT_ERR_INVALID, "NULL tree");
    for (size_t i = 0; i < t->nodes.size(); ++i) {
        if (node_keys) node_keys[i] = t->nodes[i].key;
        if (node_visits) node_visits[i] = t->nodes[i].visits;
        if (node_reward) node_reward[i] = t->nodes[i].reward;
        if (node_exhausted) node_exhausted[i] = t->nodes[i].exhausted;
    }
    for (size_t i = 0; i < t->edges.size(); ++i) {
        if (edge_parent) edge_parent[i] = t->edges[i].parent;
        if (edge_child) edge_child[i] = t->edges[i].child;
        if (edge_visits) edge_visits[i] = t->edges[i].visits;
        if (edge_reward) edge_reward[i] = t->edges[i].reward;
    }
    return CT_OK;
}

int ct_tree_key_to_string(ct_tree* t, uint64_t key, char* buf, int32_t buflen)
{
    if (!t || !buf) return tfail(CT_ERR_INVALID, "NULL argument");
    const std::string s = t->to_string(key);
    if ((int)s.size() + 1 > buflen) return tfail(CT_ERR_INVALID, "buffer too small");
    memcpy(buf, s.c_str(), s.size() + 1);
    return CT_OK;
}

/* Parse a state string; canonical != 0 also applies get_unique_state. */
int ct_tree_string_to_key(ct_tree* t, const char* s, int32_t canonical, uint64_t* key)
{
    if (!t || !s || !key) return tfail(CT_ERR_INVALID, "NULL argument");
    uint64_t k;
    if (!t->from_string(s, &k)) return tfail(CT_ERR_INVALID, "cannot parse state '%s'", s);
    *key = canonical ? t->canonical(k) : k;
    return CT_OK;
}

/* Action ids of a packed state in the reference's order; returns the count (ids nullable). */
int ct_tree_actions(ct_tree* t, uint64_t key, int32_t* ids, int32_t cap, int32_t* n_out)
{
    if (!t || !n_out) return tfail(CT_ERR_INVALID, "NULL argument");
    std::vector<int> a;
    t->actions(key, a);
    *n_out = (int32_t)a.size();
    if (ids) for (size_t i = 0; i < a.size() && (int32_t)i < cap; ++i) ids[i] = a[i];
    return CT_OK;
}

int ct_tree_do_action(ct_tree* t, uint64_t key, int32_t action, uint64_t* out)
{
    if (!t || !out) return tfail(CT_ERR_INVALID, "NULL argument");
    *out = t->step(key, action);
    return CT_OK;
}

int ct_tree_key_to_handle(ct_tree* t, uint64_t key, uint64_t* handle)
{
    if (!t || !handle) return tfail(CT_ERR_INVALID, "NULL argument");
    if (!t->to_handle(key, handle)) return tfail(CT_ERR_UNSUPPORTED, "state has interaction types other than a/i");
    return CT_OK;
}

/* Exhaustive expansion from the root in the reference's stack order (circuitree.py:565-586). */
int ct_tree_grow(ct_tree* t)
{
    NvtxRange nvtx_range("ct_tree_grow");
    if (!t) return tfail(CT_ERR_INVALID, "NULL tree");
    std::vector<std::pair<int, int>> stack;
    std::vector<int> acts;
    const int root = t->node_of(t->root_key);
    t->actions(t->root_key, acts);
    for (int a : acts) stack.emplace_back(root, a);
    while (!stack.empty()) {
        const auto [node, action] = stack.back();
        stack.pop_back();
        if (t->nodes[node].key & kTerminalBit) continue;
        const uint64_t ck = t->step(t->nodes[node].key, action);
        int child = t->node_of(ck);
        if (child < 0) {
            child = t->add_node(ck);
            t->actions(ck, acts);
            for (int a : acts) stack.emplace_back(child, a);
        }
        if (t->edge_of(node, child) < 0) t->add_edge(node, child);
    }
    return CT_OK;
}

/* Multi-GPU root-parallel merge, step 1: everything this replica has to tell the others since the
 * previous successful call, as ONE buffer of 64-bit words:
 *   [0] kPackMagic  [1] n_node  [2] n_edge  [3] n_exhausted  [4] flags (bit 0: whole tree exhausted; bits 8..: words per key W)
 *   node keys[n_node] | node d visits[n_node] | node d reward[n_node] (double bits)
 *   edge parent keys[n_edge] | edge child keys[n_edge] | edge d visits[n_edge] | edge d reward[n_edge]
 *   keys of nodes flagged exhausted[n_exhausted]
 * A key is one word (the packed state) for SimpleNetworkGrammar trees and W = flags >> 8 words for DimersGrammar
 * trees, whose keys are process-local: the interaction list itself travels (DimerSpace::to_wire).
 * Records are sparse (only what changed) and keyed by packed state: 5-component spaces cannot be
 * indexed densely.  max_depth >= 0 restricts the exchange to states with at most that many
 * interactions (+1 for terminal states); -1 exchanges everything.  If buf is NULL or cap is too
 * small nothing is committed and *n_words is the size needed.  The statistics un-counted by
 * exhaustion (reference circuitree.py:386-420) never travel as increments: the exhaustion event does,
 * and every replica replays it (see settle_exhausted). */
constexpr int64_t kPackMagic = 0x4354444c54413033ll;   // "CTDLTA03"
constexpr int64_t kPackHeader = 5;

int ct_tree_pack_deltas(ct_tree* t, int32_t max_depth, int64_t* buf, int64_t cap, int64_t* n_words)
{
    NvtxRange nvtx_range("ct_tree_pack_deltas");
    if (!t || !n_words) return tfail(CT_ERR_INVALID, "NULL argument");
    if (t->track_exhaustion) t->settle_exhausted();
    const int64_t W = t->dim ? t->dim->wire_words() : 1;
    auto wanted = [&](uint64_t key) { return max_depth < 0 || t->depth_of(key) <= max_depth; };
    int64_t nn = 0, ne = 0;
    for (const auto& nd : t->nodes)
        if ((nd.visits != nd.sync_visits || nd.reward != nd.sync_reward) && wanted(nd.key)) ++nn;
    for (const auto& e : t->edges)
        if ((e.visits != e.sync_visits || e.reward != e.sync_reward) && wanted(t->nodes[e.child].key)) ++ne;
    const int64_t nx = (int64_t)t->newly_exhausted.size();
    const int64_t need = kPackHeader + (W + 2) * nn + (2 * W + 2) * ne + W * nx;
    *n_words = need;
    if (!buf || cap < need) return CT_OK;
    auto put_key = [&](uint64_t key, int64_t* dst) -> bool {
        if (!t->dim) { *dst = (int64_t)key; return true; }
        return t->dim->to_wire(key, dst);
    };
    int64_t* nk = buf + kPackHeader; int64_t* nv = nk + W * nn; int64_t* nr = nv + nn;
    int64_t* ep = nr + nn; int64_t* ec = ep + W * ne; int64_t* ev = ec + W * ne; int64_t* er = ev + ne; int64_t* xk = er + ne;
    // keys first: nothing is committed (no sync baseline moves) if a state does not fit the wire format
    {
        int64_t a = 0, b = 0;
        bool ok = true;
        for (const auto& nd : t->nodes) {
            if ((nd.visits == nd.sync_visits && nd.reward == nd.sync_reward) || !wanted(nd.key)) continue;
            ok = put_key(nd.key, nk + W * a) && ok;
            ++a;
        }
        for (const auto& e : t->edges) {
            if ((e.visits == e.sync_visits && e.reward == e.sync_reward) || !wanted(t->nodes[e.child].key)) continue;
            ok = put_key(t->nodes[e.parent].key, ep + W * b) && put_key(t->nodes[e.child].key, ec + W * b) && ok;
            ++b;
        }
        for (int64_t x = 0; x < nx; ++x) ok = put_key(t->newly_exhausted[(size_t)x], xk + W * x) && ok;
        if (!ok) return tfail(CT_ERR_UNSUPPORTED, "a state has more interactions than the exchange format of this grammar holds");
    }
    buf[0] = kPackMagic; buf[1] = nn; buf[2] = ne; buf[3] = nx; buf[4] = (t->exhausted_all ? 1 : 0) | (W << 8);
    int64_t a = 0, b = 0;
    for (auto& nd : t->nodes) {
        if ((nd.visits == nd.sync_visits && nd.reward == nd.sync_reward) || !wanted(nd.key)) continue;
        const double dr = nd.reward - nd.sync_reward;
        nv[a] = nd.visits - nd.sync_visits; memcpy(&nr[a], &dr, 8);
        nd.sync_visits = nd.visits; nd.sync_reward = nd.reward;
        ++a;
    }
    for (auto& e : t->edges) {
        if ((e.visits == e.sync_visits && e.reward == e.sync_reward) || !wanted(t->nodes[e.child].key)) continue;
        const double dr = e.reward - e.sync_reward;
        ev[b] = e.visits - e.sync_visits; memcpy(&er[b], &dr, 8);
        e.sync_visits = e.visits; e.sync_reward = e.reward;
        ++b;
    }
    t->newly_exhausted.clear();
    return CT_OK;
}

/* Step 2: add another replica's buffer (nodes and edges are created as needed), then replay its
 * exhaustion events.  Merged statistics are the sums of all replicas' increments -- what one tree
 * would hold had it executed every path.  *tree_exhausted (nullable) is set when the root is flagged. */
int ct_tree_merge_packed(ct_tree* t, const int64_t* buf, int64_t n_words, int32_t* tree_exhausted)
{
    NvtxRange nvtx_range("ct_tree_merge_packed");
    if (!t || !buf) return tfail(CT_ERR_INVALID, "NULL argument");
    if (n_words < kPackHeader || buf[0] != kPackMagic) return tfail(CT_ERR_INVALID, "not a delta buffer");
    const int64_t nn = buf[1], ne = buf[2], nx = buf[3];
    const int64_t W = t->dim ? t->dim->wire_words() : 1;
    if ((buf[4] >> 8) != W) return tfail(CT_ERR_INVALID, "delta buffer of another grammar (%lld words per key, this tree uses %lld)",
                                        (long long)(buf[4] >> 8), (long long)W);
    if (nn < 0 || ne < 0 || nx < 0 || kPackHeader + (W + 2) * nn + (2 * W + 2) * ne + W * nx > n_words)
        return tfail(CT_ERR_INVALID, "truncated delta buffer");
    const int64_t* nk = buf + kPackHeader; const int64_t* nv = nk + W * nn; const int64_t* nr = nv + nn;
    const int64_t* ep = nr + nn; const int64_t* ec = ep + W * ne; const int64_t* ev = ec + W * ne; const int64_t* er = ev + ne;
    const int64_t* xk = er + ne;
    auto get_key = [&](const int64_t* src, uint64_t* key) -> bool {
        if (!t->dim) { *key = (uint64_t)*src; return true; }
        return t->dim->from_wire(src, key);
    };
    // decode every key before touching the tree: a malformed buffer changes nothing
    std::vector<uint64_t> keys((size_t)(nn + 2 * ne + nx));
    for (int64_t x = 0; x < nn; ++x)
        if (!get_key(nk + W * x, &keys[(size_t)x])) return tfail(CT_ERR_INVALID, "malformed state in a delta buffer");
    for (int64_t x = 0; x < ne; ++x)
        if (!get_key(ep + W * x, &keys[(size_t)(nn + 2 * x)]) || !get_key(ec + W * x, &keys[(size_t)(nn + 2 * x + 1)]))
            return tfail(CT_ERR_INVALID, "malformed state in a delta buffer");
    for (int64_t x = 0; x < nx; ++x)
        if (!get_key(xk + W * x, &keys[(size_t)(nn + 2 * ne + x)])) return tfail(CT_ERR_INVALID, "malformed state in a delta buffer");
    for (int64_t x = 0; x < nn; ++x) {
        int id = t->node_of(keys[(size_t)x]);
        if (id < 0) id = t->add_node(keys[(size_t)x]);
        double dr; memcpy(&dr, &nr[x], 8);
        Node& nd = t->nodes[id];
        nd.visits += nv[x]; nd.reward += dr;
        nd.sync_visits += nv[x]; nd.sync_reward += dr;
    }
    for (int64_t x = 0; x < ne; ++x) {
        const uint64_t pk = keys[(size_t)(nn + 2 * x)], ck = keys[(size_t)(nn + 2 * x + 1)];
        int p = t->node_of(pk), c = t->node_of(ck);
        if (p < 0) p = t->add_node(pk);
        if (c < 0) c = t->add_node(ck);
        int e = t->edge_of(p, c);
        if (e < 0) e = t->add_edge(p, c);
        double dr; memcpy(&dr, &er[x], 8);
        Edge& ed = t->edges[e];
        ed.visits += ev[x]; ed.reward += dr;
        ed.sync_visits += ev[x]; ed.sync_reward += dr;
    }
    if (t->track_exhaustion) {
        for (int64_t x = 0; x < nx; ++x) {
            const int id = t->node_of(keys[(size_t)(nn + 2 * ne + x)]);
            if (id < 0 || t->nodes[id].exhausted) continue;    // (unknown: filtered out by max_depth)
            t->mark_exhausted(id, false);
        }
        t->settle_exhausted();
    }
    if (buf[4] & 1) t->exhausted_all = true;
    if (tree_exhausted) *tree_exhausted = t->exhausted_all ? 1 : 0;
    return CT_OK;
}

}  // extern "C"
