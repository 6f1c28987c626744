// Host-side packer: the Reads CSR (sfb_reads, host pointers) -> the compact transfer layout (sfb_wire) that crosses
// PCIe on the end-to-end path.  The inverse runs on the device (sfb_expand_reads, expand.cu).
//
// Everything is computed over contiguous chunks of alignments (groups for the read lengths) by a pool of threads;
// the exception lists of the chunks are concatenated in chunk order, so they come out ascending and the result does
// not depend on the number of threads.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "sfb200.h"

namespace {

enum Which {       // order of the arrays in sfb_wire
    W_GRP_NMATES, W_MATE_NALN, W_MATE_SEQ_LEN, W_SEQ_EXC_IDX, W_SEQ_EXC_VAL, W_ALN_LEN, W_ALN_CODE, W_BINS_DICT,
    W_BINS_EXC_IDX, W_BINS_EXC_VAL, W_ALN_NODE0, W_ALN_AL_FIRST, W_ALN_AL_LAST, W_WALK_STEP, W_EXC_IDX, W_EXC_COV,
    W_WIDE_IDX, W_WIDE_NODE, W_NODE0_ANCHOR, W_NODE0_WIDE, W_COUNT
};

struct Bytes {     // exact-size, uninitialised storage
    void* p = nullptr;
    size_t n = 0;
    Bytes() = default;
    Bytes(const Bytes&) = delete;
    Bytes& operator=(const Bytes&) = delete;
    ~Bytes() { free(p); }
    bool alloc(size_t bytes) {
        free(p);
        n = bytes;
        p = bytes ? malloc(bytes) : nullptr;
        return !bytes || p;
    }
    template <class T>
    T* as() { return (T*)p; }
};

template <class F>
void run_parallel(int T, size_t n_items, size_t grain, F&& body) {
    const size_t n_chunks = (n_items + grain - 1) / grain;
    const int use = (int)std::max<size_t>(1, std::min<size_t>((size_t)T, n_chunks));
    std::atomic<size_t> next{0};
    auto work = [&] {
        for (;;) {
            const size_t c = next.fetch_add(1);
            if (c >= n_chunks) return;
            body(c, c * grain, std::min(n_items, (c + 1) * grain));
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < use; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
}

template <class T>
bool gather(Bytes& dst, const std::vector<std::vector<T>>& parts) {
    size_t n = 0;
    for (const auto& v : parts) n += v.size();
    if (!dst.alloc(n * sizeof(T))) return false;
    size_t pos = 0;
    for (const auto& v : parts) {
        if (!v.empty()) memcpy(dst.as<T>() + pos, v.data(), v.size() * sizeof(T));
        pos += v.size();
    }
    return true;
}

inline uint64_t bins_key(const uint8_t* b) {
    return (uint64_t)b[0] | (uint64_t)b[1] << 8 | (uint64_t)b[2] << 16 | (uint64_t)b[3] << 24 | (uint64_t)b[4] << 32;
}

}  // namespace

struct sfb_packed {
    Bytes arr[W_COUNT];
    int64_t seq_len_common = 0;
    int status = 0;            // 0 ok, 1 the layout does not apply to this shard, 5 out of memory
    std::string msg;
};

namespace {

void pack(sfb_packed& P, const sfb_reads& r, int T) {
    const int64_t G = r.n_groups, A = r.n_alns, R = r.n_records;
    auto give_up = [&](const char* why) { P.status = 1; P.msg = why; };
    auto oom = [&] { P.status = 5; P.msg = "out of memory"; };
    constexpr size_t kGrain = 1u << 16;

    // ---- per group: counts, read lengths --------------------------------------------------------
    if (!P.arr[W_GRP_NMATES].alloc((size_t)G) || !P.arr[W_MATE_NALN].alloc((size_t)G * 2)) return oom();
    if (G) memcpy(P.arr[W_GRP_NMATES].p, r.grp_nmates, (size_t)G);
    const size_t g_chunks = ((size_t)G + kGrain - 1) / kGrain;
    std::vector<std::vector<int64_t>> len_hist(g_chunks);
    std::atomic<bool> fits{true};
    run_parallel(T, (size_t)G, kGrain, [&](size_t c, size_t lo, size_t hi) {
        std::vector<int64_t>& h = len_hist[c];
        h.assign(65536, 0);
        uint8_t* naln = P.arr[W_MATE_NALN].as<uint8_t>();
        for (size_t g = lo; g < hi; ++g)
            for (int m = 0; m < 2; ++m) {
                const int64_t n = r.grp_aln_off[2 * g + m + 1] - r.grp_aln_off[2 * g + m];
                const int32_t len = r.mate_seq_len[2 * g + m];
                if (n < 0 || n > 255 || len < 0 || len > 0xFFFF) { fits = false; return; }
                naln[2 * g + m] = (uint8_t)n;
                if (m < (int)r.grp_nmates[g]) h[(size_t)len] += 1;
            }
    });
    if (!fits) return give_up("a mate with more than 255 alignments or a read of more than 65535 bases");
    int64_t common = 0, best = 0;
    for (int64_t len = 0; len < 65536; ++len) {
        int64_t n = 0;
        for (const auto& h : len_hist) n += h.empty() ? 0 : h[(size_t)len];
        if (n > best) { best = n; common = len; }
    }
    P.seq_len_common = common;
    {
        std::vector<std::vector<int64_t>> idx(g_chunks);
        std::vector<std::vector<int32_t>> val(g_chunks);
        run_parallel(T, (size_t)G, kGrain, [&](size_t c, size_t lo, size_t hi) {
            for (size_t g = lo; g < hi; ++g)
                for (int m = 0; m < 2; ++m) {
                    const int32_t len = r.mate_seq_len[2 * g + m];
                    const int32_t expect = m < (int)r.grp_nmates[g] ? (int32_t)common : 0;
                    if (len != expect) { idx[c].push_back((int64_t)(2 * g + m)); val[c].push_back(len); }
                }
        });
        size_t n_exc = 0;
        for (const auto& v : idx) n_exc += v.size();
        if (12 * n_exc > 4 * (size_t)G) {               // too ragged: 16 bits per mate instead
            if (!P.arr[W_MATE_SEQ_LEN].alloc((size_t)G * 2 * sizeof(uint16_t))) return oom();
            uint16_t* out = P.arr[W_MATE_SEQ_LEN].as<uint16_t>();
            run_parallel(T, (size_t)G * 2, kGrain, [&](size_t, size_t lo, size_t hi) {
                for (size_t i = lo; i < hi; ++i) out[i] = (uint16_t)r.mate_seq_len[i];
            });
        } else if (!gather(P.arr[W_SEQ_EXC_IDX], idx) || !gather(P.arr[W_SEQ_EXC_VAL], val)) {
            return oom();
        }
    }

    // ---- histogram bins: dictionary of the frequent 5-tuples (of a sample), one byte per alignment --
    std::vector<uint64_t> top;
    {
        const int64_t stride = std::max<int64_t>(1, A / 2000000);
        std::unordered_map<uint64_t, int64_t> count;
        for (int64_t a = 0; a < A; a += stride) count[bins_key(r.aln_bins + a * 5)] += 1;
        std::vector<std::pair<uint64_t, int64_t>> all(count.begin(), count.end());
        std::sort(all.begin(), all.end(), [](const auto& x, const auto& y) { return x.second != y.second ? x.second > y.second : x.first < y.first; });
        if (all.size() > 255) all.resize(255);
        for (const auto& kv : all) top.push_back(kv.first);
        std::sort(top.begin(), top.end());
    }
    if (!P.arr[W_BINS_DICT].alloc(255 * 5)) return oom();
    memset(P.arr[W_BINS_DICT].p, 0, 255 * 5);
    for (size_t i = 0; i < top.size(); ++i)
        for (int k = 0; k < 5; ++k) P.arr[W_BINS_DICT].as<uint8_t>()[i * 5 + k] = (uint8_t)(top[i] >> (8 * k));

    // ---- per alignment and per record -------------------------------------------------------------
    if (!P.arr[W_ALN_LEN].alloc((size_t)A) || !P.arr[W_ALN_CODE].alloc((size_t)A) || !P.arr[W_ALN_NODE0].alloc((size_t)A * 2) ||
        !P.arr[W_ALN_AL_FIRST].alloc((size_t)A) || !P.arr[W_ALN_AL_LAST].alloc((size_t)A) || !P.arr[W_WALK_STEP].alloc(((size_t)R + 1) / 2))
        return oom();
    if (R) memset(P.arr[W_WALK_STEP].p, 0, ((size_t)R + 1) / 2);        // nibbles are OR-ed in
    std::vector<int32_t> node0_full;                      // first nodes as they are; coded per block of 256 below
    try {
        node0_full.resize((size_t)A);
    } catch (const std::bad_alloc&) {
        return oom();
    }
    const size_t a_chunks = ((size_t)A + kGrain - 1) / kGrain;
    std::vector<std::vector<int64_t>> bins_idx(a_chunks), exc_idx(a_chunks), wide_idx(a_chunks);
    std::vector<std::vector<uint8_t>> bins_val(a_chunks);
    std::vector<std::vector<double>> exc_cov(a_chunks);
    std::vector<std::vector<int32_t>> wide_node(a_chunks);
    run_parallel(T, (size_t)A, kGrain, [&](size_t c, size_t lo, size_t hi) {
        uint8_t* aln_len = P.arr[W_ALN_LEN].as<uint8_t>();
        uint8_t* code = P.arr[W_ALN_CODE].as<uint8_t>();
        int32_t* node0 = node0_full.data();
        uint8_t* al_first = P.arr[W_ALN_AL_FIRST].as<uint8_t>();
        uint8_t* al_last = P.arr[W_ALN_AL_LAST].as<uint8_t>();
        uint8_t* step = P.arr[W_WALK_STEP].as<uint8_t>();
        // one nibble per record: the two bytes at the ends of this chunk's record range may be shared with the
        // neighbouring chunks (another thread), so nibbles landing there are OR-ed in atomically
        const int64_t first_byte = r.aln_walk_off[lo] >> 1, last_byte = (r.aln_walk_off[hi] - 1) >> 1;
        auto put = [&](int64_t q, unsigned nib) {
            const int64_t b = q >> 1;
            const uint8_t v = (uint8_t)(nib << (4 * (q & 1)));
            if (b == first_byte || b == last_byte) __atomic_fetch_or(step + b, v, __ATOMIC_RELAXED);
            else step[b] |= v;
        };
        for (size_t a = lo; a < hi; ++a) {
            const int64_t r0 = r.aln_walk_off[a], r1 = r.aln_walk_off[a + 1];
            if (r1 - r0 < 0 || r1 - r0 > 255) { fits = false; return; }
            aln_len[a] = (uint8_t)(r1 - r0);
            const uint64_t key = bins_key(r.aln_bins + a * 5);
            const auto it = std::lower_bound(top.begin(), top.end(), key);
            if (it != top.end() && *it == key) {
                code[a] = (uint8_t)(it - top.begin());
            } else {
                code[a] = 255;
                bins_idx[c].push_back((int64_t)a);
                bins_val[c].insert(bins_val[c].end(), r.aln_bins + a * 5, r.aln_bins + a * 5 + 5);
            }
            node0[a] = r1 > r0 ? r.walk_node[r0] : 0;
            al_first[a] = al_last[a] = 0;
            for (int64_t q = r0; q < r1; ++q) {
                // node id: the step from the previous record of the walk
                if (q > r0) {
                    const int64_t d = (int64_t)r.walk_node[q] - (int64_t)r.walk_node[q - 1];
                    if (d < -7 || d > 7) {                      // nibble 0: listed
                        wide_idx[c].push_back(q);
                        wide_node[c].push_back(r.walk_node[q]);
                    } else {
                        put(q, (unsigned)(d + 8));
                    }
                }
                // aligned bases: a byte at the two ends of the walk, the node length in between, else explicit
                const int32_t v = r.walk_alen[q];
                const int32_t low = v & 0xFFFF, nl = v >> 16;
                const bool edge = q == r0 || q == r1 - 1;
                const bool listed = v < 0 || (edge ? low >= 255 : low != nl);
                if (listed) {
                    exc_idx[c].push_back(q);
                    exc_cov[c].push_back(v < 0 ? r.exc_cov[-(int64_t)v - 1] : (double)low / (double)std::max(nl, 1));
                }
                if (q == r0) al_first[a] = listed ? 255 : (uint8_t)low;
                else if (q == r1 - 1) al_last[a] = listed ? 255 : (uint8_t)low;
            }
        }
    });
    if (!fits) return give_up("a walk of more than 255 nodes");
    // ---- first nodes: 16 bits above the smallest one of the block of 256 alignments (blocks that span more: listed) ----
    {
        const size_t nb = ((size_t)A + 255) / 256;
        if (!P.arr[W_NODE0_ANCHOR].alloc(nb * 4)) return oom();
        int32_t* anchor = P.arr[W_NODE0_ANCHOR].as<int32_t>();
        uint16_t* lo16 = P.arr[W_ALN_NODE0].as<uint16_t>();
        run_parallel(T, nb, 256, [&](size_t, size_t b0, size_t b1) {
            for (size_t b = b0; b < b1; ++b) {
                const size_t a0 = b * 256, a1 = std::min((size_t)A, a0 + 256);
                int32_t mn = node0_full[a0], mx = node0_full[a0];
                for (size_t a = a0; a < a1; ++a) {
                    mn = std::min(mn, node0_full[a]);
                    mx = std::max(mx, node0_full[a]);
                }
                const bool narrow = mn >= 0 && (int64_t)mx - (int64_t)mn <= 65535;
                anchor[b] = narrow ? mn : -1;
                for (size_t a = a0; a < a1; ++a) lo16[a] = narrow ? (uint16_t)(node0_full[a] - mn) : (uint16_t)0;
            }
        });
        size_t n_wide_blocks = 0;
        for (size_t b = 0; b < nb; ++b)
            if (anchor[b] < 0) anchor[b] = -1 - (int32_t)n_wide_blocks++;
        if (!P.arr[W_NODE0_WIDE].alloc(n_wide_blocks * 256 * 4)) return oom();
        int32_t* wide0 = P.arr[W_NODE0_WIDE].as<int32_t>();
        run_parallel(T, nb, 256, [&](size_t, size_t b0, size_t b1) {
            for (size_t b = b0; b < b1; ++b) {
                if (anchor[b] >= 0) continue;
                const size_t k = (size_t)(-1 - anchor[b]), a0 = b * 256;
                for (size_t j = 0; j < 256; ++j) wide0[256 * k + j] = a0 + j < (size_t)A ? node0_full[a0 + j] : 0;
            }
        });
    }
    size_t n_exc = 0, n_wide = 0;
    for (const auto& v : exc_idx) n_exc += v.size();
    for (const auto& v : wide_idx) n_wide += v.size();
    if (16 * n_exc + 12 * n_wide > 2 * (size_t)R) return give_up("too many explicit coverages / wide node steps");
    if (!gather(P.arr[W_BINS_EXC_IDX], bins_idx) || !gather(P.arr[W_BINS_EXC_VAL], bins_val) || !gather(P.arr[W_EXC_IDX], exc_idx) ||
        !gather(P.arr[W_EXC_COV], exc_cov) || !gather(P.arr[W_WIDE_IDX], wide_idx) || !gather(P.arr[W_WIDE_NODE], wide_node))
        return oom();
}

}  // namespace

extern "C" {

sfb_packed* sfb_wire_pack(const sfb_reads* host_reads, int n_threads) {
    sfb_packed* P = new sfb_packed();
    try {
        pack(*P, *host_reads, std::max(1, std::min(n_threads, 4096)));
    } catch (const std::exception& ex) {
        P->status = 5;
        P->msg = ex.what();
    }
    return P;
}

int sfb_packed_status(const sfb_packed* P, const char** msg) {
    if (msg) *msg = P->msg.c_str();
    return P->status;
}

int64_t sfb_packed_size(const sfb_packed* P, int which) {
    if (which == 100) return P->seq_len_common;
    return which >= 0 && which < W_COUNT ? (int64_t)P->arr[which].n : -1;
}

int sfb_packed_copy(const sfb_packed* P, int which, void* dst) {
    if (which < 0 || which >= W_COUNT) return -1;
    if (P->arr[which].n) memcpy(dst, P->arr[which].p, P->arr[which].n);
    return 0;
}

void sfb_packed_free(sfb_packed* P) { delete P; }

}  // extern "C"
