// Host-side locality order of the read groups and the permuted Reads CSR (no GPU involved).
//
// The abundance pass is a chain of gathers into the graph index, the paths and the per-path / per-slot accumulators.
// Its sums do not depend on the order of the read groups, so the host may stage them in the order in which
// neighbouring groups touch the same part of the graph: stable by the first node of the group's first retained
// alignment (node ids of a pangenome built cluster by cluster are contiguous per cluster), groups without an
// alignment last.  Same result as schema.Reads.locality_order() + permute_groups() in numpy, multi-threaded:
// keys -> stable LSD radix sort of the group indices (16-bit digits, per-thread histograms) -> prefix sums ->
// parallel copy of the mates, alignments and records of every group.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "sfb200.h"

namespace {

template <class F>
void for_chunks(int T, size_t n_items, F&& body) {      // chunk c of exactly T contiguous chunks (some may be empty)
    std::vector<std::thread> pool;
    auto work = [&](int c) {
        const size_t lo = n_items * (size_t)c / (size_t)T, hi = n_items * (size_t)(c + 1) / (size_t)T;
        body(c, lo, hi);
    };
    for (int c = 1; c < T; ++c) pool.emplace_back(work, c);
    work(0);
    for (auto& th : pool) th.join();
}


// Stable order of the groups by key (keys in [0, none]) and the permuted CSR; shared by the two entry points.
int permute_by_key(const sfb_reads* r, int T, const std::vector<uint32_t>& key, uint32_t none, int64_t* order,
                   int64_t* aln_index, const sfb_reads* dst) {
    const int64_t G = r->n_groups, A = r->n_alns;
    uint8_t* d_nmates = const_cast<uint8_t*>(dst->grp_nmates);
    int64_t* d_grp_off = const_cast<int64_t*>(dst->grp_aln_off);
    int32_t* d_seq_len = const_cast<int32_t*>(dst->mate_seq_len);
    int64_t* d_walk_off = const_cast<int64_t*>(dst->aln_walk_off);
    uint8_t* d_bins = const_cast<uint8_t*>(dst->aln_bins);
    int32_t* d_node = const_cast<int32_t*>(dst->walk_node);
    int32_t* d_alen = const_cast<int32_t*>(dst->walk_alen);
    {
        // ---- stable LSD radix sort of the group indices by key, 16 bits per pass ----
        std::vector<int64_t> idx_a((size_t)G), idx_b((size_t)G);
        for_chunks(T, (size_t)G, [&](int, size_t lo, size_t hi) {
            for (size_t g = lo; g < hi; ++g) idx_a[g] = (int64_t)g;
        });
        int64_t* src = idx_a.data();
        int64_t* out = idx_b.data();
        const int passes = none >> 16 ? 2 : 1;
        std::vector<std::vector<size_t>> hist((size_t)T, std::vector<size_t>(65536));
        for (int pass = 0; pass < passes; ++pass) {
            const int shift = 16 * pass;
            for_chunks(T, (size_t)G, [&](int c, size_t lo, size_t hi) {
                auto& h = hist[c];
                std::fill(h.begin(), h.end(), 0);
                for (size_t i = lo; i < hi; ++i) h[(key[src[i]] >> shift) & 0xffff] += 1;
            });
            size_t pos = 0;                                  // digit-major, chunk-minor: stable
            for (size_t dgt = 0; dgt < 65536; ++dgt)
                for (int c = 0; c < T; ++c) {
                    const size_t n = hist[c][dgt];
                    hist[c][dgt] = pos;
                    pos += n;
                }
            for_chunks(T, (size_t)G, [&](int c, size_t lo, size_t hi) {
                auto& h = hist[c];
                for (size_t i = lo; i < hi; ++i) out[h[(key[src[i]] >> shift) & 0xffff]++] = src[i];
            });
            std::swap(src, out);
        }
        memcpy(order, src, (size_t)G * sizeof(int64_t));
        // ---- offsets of the permuted arrays (serial prefix sums: one pass over the groups, one over the alignments) ----
        d_grp_off[0] = 0;
        for (int64_t k = 0; k < G; ++k) {
            const int64_t g = order[k];
            d_grp_off[2 * k + 1] = d_grp_off[2 * k] + (r->grp_aln_off[2 * g + 1] - r->grp_aln_off[2 * g]);
            d_grp_off[2 * k + 2] = d_grp_off[2 * k + 1] + (r->grp_aln_off[2 * g + 2] - r->grp_aln_off[2 * g + 1]);
        }
        for_chunks(T, (size_t)G, [&](int, size_t lo, size_t hi) {
            for (size_t k = lo; k < hi; ++k) {
                const int64_t g = order[k];
                d_nmates[k] = r->grp_nmates[g];
                d_seq_len[2 * k] = r->mate_seq_len[2 * g];
                d_seq_len[2 * k + 1] = r->mate_seq_len[2 * g + 1];
                const int64_t n = r->grp_aln_off[2 * g + 2] - r->grp_aln_off[2 * g];
                for (int64_t j = 0; j < n; ++j) aln_index[d_grp_off[2 * k] + j] = r->grp_aln_off[2 * g] + j;
            }
        });
        if (A > 0 || d_walk_off) d_walk_off[0] = 0;
        for (int64_t a = 0; a < A; ++a) {
            const int64_t old = aln_index[a];
            d_walk_off[a + 1] = d_walk_off[a] + (r->aln_walk_off[old + 1] - r->aln_walk_off[old]);
        }
        for_chunks(T, (size_t)A, [&](int, size_t lo, size_t hi) {
            for (size_t a = lo; a < hi; ++a) {
                const int64_t old = aln_index[a];
                memcpy(d_bins + 5 * a, r->aln_bins + 5 * old, 5);
                const int64_t w = r->aln_walk_off[old], n = r->aln_walk_off[old + 1] - w;
                if (n > 0) {
                    memcpy(d_node + d_walk_off[a], r->walk_node + w, (size_t)n * sizeof(int32_t));
                    memcpy(d_alen + d_walk_off[a], r->walk_alen + w, (size_t)n * sizeof(int32_t));
                }
            }
        });
    }
    return SFB_OK;
}

}  // namespace

extern "C" int sfb_locality_permute(const sfb_reads* r, int n_threads, int64_t* order, int64_t* aln_index,
                                    const sfb_reads* dst) {
    if (!r || !dst || !order || (r->n_alns && !aln_index)) return SFB_E_ARG;
    const int64_t G = r->n_groups, A = r->n_alns, R = r->n_records;
    if (G < 0 || A < 0 || R < 0) return SFB_E_ARG;
    const int T = std::max(1, std::min(n_threads, 64));
    try {
        // ---- keys: first node of the first retained alignment; none -> after every node ----
        std::vector<uint32_t> key((size_t)G);
        std::vector<uint32_t> max_part((size_t)T, 0);
        for_chunks(T, (size_t)G, [&](int c, size_t lo, size_t hi) {
            uint32_t mx = 0;
            for (size_t g = lo; g < hi; ++g) {
                const int64_t a0 = r->grp_aln_off[2 * g], a1 = r->grp_aln_off[2 * g + 2];
                uint32_t k = 0xffffffffu;
                if (a1 > a0 && R > 0) {
                    const int64_t w = std::min<int64_t>(r->aln_walk_off[a0], R - 1);
                    k = (uint32_t)r->walk_node[w];
                    mx = std::max(mx, k);
                }
                key[g] = k;
            }
            max_part[c] = mx;
        });
        uint32_t top = 0;
        for (uint32_t v : max_part) top = std::max(top, v);
        const uint32_t none = top + 1;                     // groups without an alignment sort last
        for_chunks(T, (size_t)G, [&](int, size_t lo, size_t hi) {
            for (size_t g = lo; g < hi; ++g)
                if (key[g] == 0xffffffffu) key[g] = none;
        });
        return permute_by_key(r, T, key, none, order, aln_index, dst);
    } catch (const std::exception&) {
        return SFB_E_ARG;                                    // out of memory
    }
}

// Tile order (sfb_tiles in sfb200.h): the groups of a tile -- the tile that holds the first and the last node of every
// alignment of the group -- together, tiles in order, inside a tile by the first node of the group's first alignment;
// groups without such a tile (mates in different tiles, a node outside every tile, no alignment) go last.
extern "C" int sfb_tile_permute(const sfb_reads* r, const int32_t* node_tile, int64_t n_tiles, int n_threads,
                                int64_t* order, int64_t* aln_index, const sfb_reads* dst, int64_t* tile_grp_off) {
    if (!r || !dst || !order || !tile_grp_off || (r->n_alns && (!aln_index || !node_tile))) return SFB_E_ARG;
    const int64_t G = r->n_groups, A = r->n_alns, R = r->n_records;
    if (G < 0 || A < 0 || R < 0 || n_tiles < 0 || n_tiles >= 0x7fffffffLL) return SFB_E_ARG;
    const int T = std::max(1, std::min(n_threads, 64));
    try {
        // Inside a tile the groups are ordered by the first node of their first alignment: tiles are built in node order
        // (tile t's nodes lie before tile t+1's), so ONE key -- that node, or "none" -- orders the groups by tile and,
        // within a tile, by position: neighbouring alignments then gather the same index records, path constants and
        // accumulator lines.
        std::vector<uint32_t> key((size_t)G);
        std::vector<std::vector<int64_t>> count((size_t)T);
        std::vector<uint32_t> max_part((size_t)T, 0);
        for_chunks(T, (size_t)G, [&](int c, size_t lo, size_t hi) {
            auto& cnt = count[c];
            cnt.assign((size_t)n_tiles + 1, 0);
            uint32_t mx = 0;
            for (size_t g = lo; g < hi; ++g) {
                const int64_t a0 = r->grp_aln_off[2 * g], a1 = r->grp_aln_off[2 * g + 2];
                int64_t tile = -1;
                uint32_t node = 0xffffffffu;
                const int64_t am = r->grp_aln_off[2 * g + 1];
                bool ok = a1 > a0 && am - a0 <= 255 && a1 - am <= 255;      // the kernels' chunks hold 2 * 255 alignments
                for (int64_t a = a0; a < a1 && ok; ++a) {
                    const int64_t w0 = r->aln_walk_off[a], w1 = r->aln_walk_off[a + 1];
                    if (w1 <= w0) continue;                               // an empty walk matches nothing anywhere
                    const int32_t t0 = node_tile[r->walk_node[w0]], t1 = node_tile[r->walk_node[w1 - 1]];
                    if (t0 < 0 || t0 != t1 || (tile >= 0 && t0 != tile)) ok = false;
                    if (tile < 0) node = (uint32_t)r->walk_node[w0];
                    tile = t0;
                }
                const bool in = ok && tile >= 0;
                key[g] = in ? node : 0xffffffffu;
                if (in) mx = std::max(mx, node);
                cnt[in ? (size_t)tile : (size_t)n_tiles] += 1;
            }
            max_part[c] = mx;
        });
        uint32_t top = 0;
        for (uint32_t v : max_part) top = std::max(top, v);
        const uint32_t none = top + 1;                     // the groups outside the tiles sort last, in file order
        for_chunks(T, (size_t)G, [&](int, size_t lo, size_t hi) {
            for (size_t g = lo; g < hi; ++g)
                if (key[g] == 0xffffffffu) key[g] = none;
        });
        tile_grp_off[0] = 0;
        for (int64_t t = 0; t <= n_tiles; ++t) {
            int64_t n = 0;
            for (int c = 0; c < T; ++c) n += count[c][(size_t)t];
            tile_grp_off[t + 1] = tile_grp_off[t] + n;
        }
        return permute_by_key(r, T, key, none, order, aln_index, dst);
    } catch (const std::exception&) {
        return SFB_E_ARG;                                    // out of memory
    }
}
