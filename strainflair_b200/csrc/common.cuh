// Shared device helpers for libsfb200 (sm_100a).  Internal; the public surface is include/sfb200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sfb200.h"

namespace sfb {

typedef long long i64;
typedef unsigned long long u64;

int fail(int code, const char* fmt, ...);
int check_launch(const char* what);

#define SFB_REQUIRE(cond, msg)                                        \
    do {                                                              \
        if (!(cond)) return ::sfb::fail(SFB_E_ARG, "%s: %s", __func__, msg); \
    } while (0)

constexpr int kWarp = 32;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- match list accessors (encoding in sfb200.h) ---------------------------
__device__ __forceinline__ int match_count(const sfb_work& w, i64 a) {
    const uint32_t v = w.aln_match[a];
    if (v == SFB_MATCH_NONE) return 0;
    if (!(v & SFB_MATCH_MULTI)) return 1;
    return w.match_buf[(size_t)(v & 0x7fffffffu) * 2];
}
__device__ __forceinline__ uint32_t match_get(const sfb_work& w, i64 a, int k) {
    const uint32_t v = w.aln_match[a];
    if (!(v & SFB_MATCH_MULTI)) return v;
    return (uint32_t)w.match_buf[(size_t)(v & 0x7fffffffu) * 2 + 1 + k];
}

// Coverage of walk record `rec` (position i of an L-node walk matched at `code`):
// aligned bases / node length, the node being the path node the walk position maps to
// (the reversed walk reads the path backwards), json2csv.py:574.
__device__ __forceinline__ double record_cov(const sfb_graph& g, const sfb_reads& r, i64 rec, uint32_t code,
                                             int i, int L) {
    const int alen = r.walk_alen[rec];
    if (alen < 0) return r.exc_cov[-(i64)alen - 1];
    const uint32_t slot0 = code & SFB_SLOT_MASK;
    const uint32_t s = (code & SFB_MATCH_REV) ? slot0 + (uint32_t)(L - 1 - i) : slot0 + (uint32_t)i;
    return (double)alen / (double)g.slot_len[s];
}

// A mate of a read group: its retained alignments [a0, a0+na) and the total number of
// (path, position) match tuples over them ("colored_match", json2csv.py:851-866).
struct Mate {
    i64 a0;
    int na;
    int nt;
    int seq_len;
    int m;   // mate slot inside the group
};

__device__ __forceinline__ void mate_load(const sfb_reads& r, const sfb_work& w, i64 g, int m, Mate& out) {
    out.m = m;
    out.a0 = r.grp_aln_off[2 * g + m];
    out.na = (int)(r.grp_aln_off[2 * g + m + 1] - out.a0);
    out.seq_len = r.mate_seq_len[2 * g + m];
    out.nt = 0;
}
__device__ __forceinline__ void mate_count_tuples(const sfb_work& w, Mate& mt) {
    int n = 0;
    for (int j = 0; j < mt.na; ++j) n += match_count(w, mt.a0 + j);
    mt.nt = n;
}

// Visit tuples in colored_match order: f(index, code, path id, local alignment index); stop when f returns false.
template <class F>
__device__ __forceinline__ void for_tuples(const sfb_graph& g, const sfb_work& w, const Mate& mt, F f) {
    int idx = 0;
    for (int j = 0; j < mt.na; ++j) {
        const int c = match_count(w, mt.a0 + j);
        for (int k = 0; k < c; ++k, ++idx) {
            const uint32_t code = match_get(w, mt.a0 + j, k);
            if (!f(idx, code, g.slot_path[code & SFB_SLOT_MASK], j)) return;
        }
    }
}
// list_path_id.count(pid) (json2csv.py:932)
__device__ __forceinline__ int count_pid(const sfb_graph& g, const sfb_work& w, const Mate& mt, int pid) {
    int n = 0;
    for_tuples(g, w, mt, [&](int, uint32_t, int p, int) { n += (p == pid); return true; });
    return n;
}
// true iff no tuple before index i carries pid
__device__ __forceinline__ bool first_of_pid(const sfb_graph& g, const sfb_work& w, const Mate& mt, int i, int pid) {
    bool first = true;
    for_tuples(g, w, mt, [&](int j, uint32_t, int p, int) {
        if (j >= i) return false;
        if (p == pid) { first = false; return false; }
        return true;
    });
    return first;
}

// Strain-set algebra on the per-path bitmasks, one 64-bit word at a time.
__device__ __forceinline__ u64 mate_strain_union(const sfb_graph& g, const sfb_work& w, const Mate& mt, int word) {
    u64 u = 0;
    for_tuples(g, w, mt, [&](int, uint32_t, int p, int) { u |= g.path_smask[(size_t)p * g.mask_words + word]; return true; });
    return u;
}

struct BlockCounters {
    int v[SFB_N_COUNTERS];
};
__device__ __forceinline__ void block_counters_zero(BlockCounters& sc) {
    for (int i = threadIdx.x; i < SFB_N_COUNTERS; i += blockDim.x) sc.v[i] = 0;
}
__device__ __forceinline__ void block_counters_flush(BlockCounters& sc, long long* counters) {
    for (int i = threadIdx.x; i < SFB_N_COUNTERS; i += blockDim.x)
        if (sc.v[i]) atomicAdd((u64*)&counters[i], (u64)(long long)sc.v[i]);
}
#define SFB_COUNT(sc, which, n) atomicAdd(&(sc).v[which], (n))

}  // namespace sfb
