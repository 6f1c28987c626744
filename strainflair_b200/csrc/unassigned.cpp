// Cluster-level analysis of the unassigned reads: Cluster.min_strains_inference (json2csv.py:149-201) with
// canonical_tiny (:63-71), check_incomp_comp (:83-97) and sublist (:99-100), for all clusters at once.
//
// The reference builds, per cluster, a graph over the unassigned walks (edge = the two walks share a node but
// disagree around it), takes the size of its largest clique, filters the walks by the support their compatible
// walks give them (scipy.signal.find_peaks on the negated support profile), and takes the largest clique again.
// Two details of the Python runtime decide results and are restated here:
//   * check_incomp_comp anchors the comparison at `list(set(a) & set(b))[0]`: the first element, in hash-table
//     order, of a CPython set intersection.  PySet below replays CPython's open addressing (setobject.c: table of
//     8 growing x4, linear probes of 9 then the perturbed sequence, hash(int) = int) so that the same node is picked.
//   * find_peaks(-s) reports a plateau as a peak only when it rises before and falls after inside the array.
// Largest cliques are order independent, so a branch-and-bound search replaces networkx.find_cliques.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "sfb200.h"

namespace {

// ---- CPython's set of ints (Objects/setobject.c, 3.8 - 3.12), insert-only ----------------------------
struct PySet {
    static constexpr size_t kLinear = 9;
    std::vector<int64_t> key;
    std::vector<uint8_t> used;
    size_t mask = 7, fill = 0;
    PySet() : key(8, 0), used(8, 0) {}

    static uint64_t hash_of(int64_t v) {
        // hash(int): the value modulo 2^61 - 1 with the sign kept, -1 mapped to -2
        constexpr int64_t M = (int64_t(1) << 61) - 1;
        int64_t h = v >= 0 ? v % M : -((-v) % M);
        if (h == -1) h = -2;
        return (uint64_t)h;
    }
    static void insert_clean(std::vector<int64_t>& k, std::vector<uint8_t>& u, size_t mask, int64_t v) {
        const uint64_t hash = hash_of(v);
        uint64_t perturb = hash;
        size_t i = (size_t)hash & mask;
        for (;;) {
            if (!u[i]) { k[i] = v; u[i] = 1; return; }
            if (i + kLinear <= mask)
                for (size_t j = 1; j <= kLinear; ++j)
                    if (!u[i + j]) { k[i + j] = v; u[i + j] = 1; return; }
            perturb >>= 5;
            i = (i * 5 + 1 + (size_t)perturb) & mask;
        }
    }
    void resize(size_t minused) {
        size_t n = 8;
        while (n <= minused) n <<= 1;
        std::vector<int64_t> k(n, 0);
        std::vector<uint8_t> u(n, 0);
        for (size_t i = 0; i <= mask; ++i)
            if (used[i]) insert_clean(k, u, n - 1, key[i]);
        key.swap(k);
        used.swap(u);
        mask = n - 1;
    }
    bool contains(int64_t v) const {
        const uint64_t hash = hash_of(v);
        uint64_t perturb = hash;
        size_t i = (size_t)hash & mask;
        for (;;) {
            size_t probes = (i + kLinear <= mask) ? kLinear : 0;
            for (size_t j = 0; j <= probes; ++j) {
                if (!used[i + j]) return false;
                if (key[i + j] == v) return true;
            }
            perturb >>= 5;
            i = (i * 5 + 1 + (size_t)perturb) & mask;
        }
    }
    void add(int64_t v) {
        const uint64_t hash = hash_of(v);
        uint64_t perturb = hash;
        size_t i = (size_t)hash & mask;
        for (;;) {
            size_t probes = (i + kLinear <= mask) ? kLinear : 0;
            for (size_t j = 0; j <= probes; ++j) {
                if (!used[i + j]) {
                    key[i + j] = v;
                    used[i + j] = 1;
                    ++fill;
                    if (fill * 5 >= mask * 3) resize(fill > 50000 ? fill * 2 : fill * 4);
                    return;
                }
                if (key[i + j] == v) return;
            }
            perturb >>= 5;
            i = (i * 5 + 1 + (size_t)perturb) & mask;
        }
    }
    size_t size() const { return fill; }
};

// first element of list(set(a) & set(b)), false when the sets are disjoint (set_intersection: the smaller set is
// walked in table order -- `b`'s when the sizes are equal -- and its members found in the other go to a new set)
bool first_shared(const PySet& a, const PySet& b, int64_t& pivot) {
    const PySet& walk = b.size() > a.size() ? a : b;
    const PySet& other = b.size() > a.size() ? b : a;
    PySet out;
    for (size_t i = 0; i <= walk.mask; ++i)
        if (walk.used[i] && other.contains(walk.key[i])) out.add(walk.key[i]);
    if (out.size() == 0) return false;
    for (size_t i = 0; i <= out.mask; ++i)
        if (out.used[i]) { pivot = out.key[i]; return true; }
    return false;
}

struct Walk {
    std::vector<int64_t> nodes;      // oriented (canonical_tiny)
    std::vector<int64_t> sorted;     // distinct nodes, ascending
    PySet set;
};

// 1: the walks agree around a shared node, -1: they share a node but disagree, 0: disjoint (json2csv.py:83-97)
int relation(const Walk& a, const Walk& b) {
    int64_t pivot;
    if (!first_shared(a.set, b.set, pivot)) return 0;
    const std::vector<int64_t>&x = a.nodes, &y = b.nodes;
    for (size_t i = 0; i < x.size(); ++i) {
        if (x[i] != pivot) continue;
        for (size_t j = 0; j < y.size(); ++j) {
            if (y[j] != pivot) continue;
            const size_t left = std::min(i, j), right = std::min(x.size() - i, y.size() - j);
            if (std::equal(x.begin() + (i - left), x.begin() + (i + right), y.begin() + (j - left))) return 1;
        }
    }
    return -1;
}

inline bool subset(const std::vector<int64_t>& a, const std::vector<int64_t>& b) {      // a <= b, both sorted, distinct
    return std::includes(b.begin(), b.end(), a.begin(), a.end());
}

// scipy.signal._peak_finding_utils._local_maxima_1d: is there any peak in x?
bool has_peak(const std::vector<int64_t>& x) {
    const size_t n = x.size();
    if (n < 3) return false;
    const size_t i_max = n - 1;
    for (size_t i = 1; i < i_max; ++i) {
        if (x[i - 1] < x[i]) {
            size_t ahead = i + 1;
            while (ahead < i_max && x[ahead] == x[i]) ++ahead;
            if (x[ahead] < x[i]) return true;
        }
    }
    return false;
}

// ---- largest clique of a small graph (adjacency as bit rows) ------------------------------------------
struct Clique {
    size_t words;
    const std::vector<uint64_t>* adj;
    int best = 0;
    static int count(const std::vector<uint64_t>& s) {
        int c = 0;
        for (uint64_t w : s) c += __builtin_popcountll(w);
        return c;
    }
    void grow(std::vector<uint64_t>& cand, int size) {
        // Bron-Kerbosch without the excluded set (only the size is wanted), pruned by size + |candidates|
        for (;;) {
            const int left = count(cand);
            if (size + left <= best) return;
            if (left == 0) { best = size; return; }
            size_t v = 0;
            for (size_t w = 0; w < words; ++w)
                if (cand[w]) { v = w * 64 + (size_t)__builtin_ctzll(cand[w]); break; }
            std::vector<uint64_t> next(words);
            for (size_t w = 0; w < words; ++w) next[w] = cand[w] & (*adj)[v * words + w];
            grow(next, size + 1);
            cand[v / 64] &= ~(uint64_t(1) << (v % 64));        // then the cliques without v
        }
    }
    int largest(const std::vector<uint64_t>& nodes) {
        best = 0;
        std::vector<uint64_t> cand = nodes;
        grow(cand, 0);
        return best;
    }
};

void one_cluster(const int64_t* walk_off, const int64_t* nodes, int64_t w0, int64_t w1, int64_t& before, int64_t& kept_n,
                 int64_t& after, double& support) {
    const size_t n = (size_t)(w1 - w0);
    std::vector<Walk> walks(n);
    for (size_t k = 0; k < n; ++k) {
        Walk& w = walks[k];
        w.nodes.assign(nodes + walk_off[w0 + (int64_t)k], nodes + walk_off[w0 + (int64_t)k + 1]);
        if (w.nodes.size() > 1) {               // canonical_tiny: mostly descending ids -> reversed
            size_t rising = 0;
            for (size_t i = 0; i + 1 < w.nodes.size(); ++i) rising += w.nodes[i + 1] - w.nodes[i] > 0;
            if (!(2 * rising > w.nodes.size() - 1)) std::reverse(w.nodes.begin(), w.nodes.end());
        }
        for (int64_t v : w.nodes) w.set.add(v);
        w.sorted = w.nodes;
        std::sort(w.sorted.begin(), w.sorted.end());
        w.sorted.erase(std::unique(w.sorted.begin(), w.sorted.end()), w.sorted.end());
    }
    const size_t words = (n + 63) / 64;
    std::vector<uint64_t> conflict(n * words, 0), all(words, 0);
    std::vector<std::vector<uint32_t>> friends(n);
    for (size_t i = 0; i < n; ++i) {
        all[i / 64] |= uint64_t(1) << (i % 64);
        for (size_t j = i + 1; j < n; ++j) {
            const int rel = relation(walks[i], walks[j]);
            if (rel < 0) {
                conflict[i * words + j / 64] |= uint64_t(1) << (j % 64);
                conflict[j * words + i / 64] |= uint64_t(1) << (i % 64);
            } else if (rel > 0) {
                friends[i].push_back((uint32_t)j);
                friends[j].push_back((uint32_t)i);
            }
        }
    }
    Clique cq{words, &conflict};
    before = cq.largest(all);
    std::vector<uint64_t> keep(words, 0);
    kept_n = 0;
    std::vector<int64_t> tally;
    for (size_t i = 0; i < n; ++i) {
        const Walk& me = walks[i];
        size_t proper = 0;
        for (uint32_t r : friends[i]) proper += !(subset(me.sorted, walks[r].sorted) || subset(walks[r].sorted, me.sorted));
        if (proper < 2) continue;
        std::vector<int64_t> profile(me.nodes.size(), -1);           // the NEGATED support, as find_peaks sees it
        for (uint32_t r : friends[i])
            for (int64_t v : walks[r].nodes) {
                const auto it = std::find(me.nodes.begin(), me.nodes.end(), v);     // list.index: first occurrence
                if (it != me.nodes.end()) profile[(size_t)(it - me.nodes.begin())] -= 1;
            }
        if (has_peak(profile)) continue;
        keep[i / 64] |= uint64_t(1) << (i % 64);
        kept_n += 1;
        tally.insert(tally.end(), me.nodes.begin(), me.nodes.end());
    }
    after = cq.largest(keep);
    if (tally.empty()) {
        support = std::numeric_limits<double>::quiet_NaN();       // np.mean([]) (json2csv.py:201)
    } else {
        const double total = (double)tally.size();
        std::sort(tally.begin(), tally.end());
        const double distinct = (double)(std::unique(tally.begin(), tally.end()) - tally.begin());
        support = total / distinct;
    }
}

}  // namespace

extern "C" int sfb_min_strains(int64_t n_clusters, const int64_t* clu_off, const int64_t* walk_off, const int64_t* nodes,
                               int n_threads, int64_t* before, int64_t* kept, int64_t* after, double* support) {
    if (n_clusters < 0 || (n_clusters && (!clu_off || !before || !kept || !after || !support))) return SFB_E_ARG;
    std::atomic<int64_t> next{0};
    std::atomic<bool> failed{false};
    auto work = [&] {
        try {
            for (;;) {
                const int64_t c = next.fetch_add(1);
                if (c >= n_clusters) return;
                before[c] = kept[c] = after[c] = 0;
                support[c] = 0.0;
                if (clu_off[c + 1] > clu_off[c])
                    one_cluster(walk_off, nodes, clu_off[c], clu_off[c + 1], before[c], kept[c], after[c], support[c]);
            }
        } catch (...) {
            failed = true;
        }
    };
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(n_threads, 4096), n_clusters));
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    return failed ? SFB_E_ARG : SFB_OK;
}
