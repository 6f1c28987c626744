// Register-resident fast path of the per-group logic (k_classify / k_pass2_resolve).
//
// Almost every read group has at most two retained alignments per mate and at most eight match tuples per
// mate.  For those, the tuples live in registers (fully unrolled, compile-time indices), the set logic of
// json2csv.py:922-1000 / :1127-1217 becomes 8-bit masks built from a fixed 8x8 compare matrix, and the strain
// algebra is AND/OR/popcount on one 64-bit word per path.  No thread-local memory, no data-dependent loops:
// lanes of a warp stay converged whatever their tuple counts are.  Groups that do not fit (or graphs with more
// than 64 strains) take the generic LocalView/GlobalView code.
#pragma once
#include "common.cuh"

namespace sfb {

#ifndef SFB_KR
#define SFB_KR 8
#endif
constexpr int kR = SFB_KR;

struct RegMate {
    int n;            // tuples
    int c0;           // tuples of the first alignment (the rest belong to the second)
    int pid[kR];
    uint32_t code[kR];
};

// `sentinel`: distinct negative ids for the unused entries, different for the two mates, so that they never compare equal
__device__ __forceinline__ bool reg_load(const sfb_work& w, const Mate& mt, int sentinel, RegMate& r) {
    if (mt.na > 2) return false;
    uint2 ma = make_uint2(SFB_MATCH_NONE, 0u), mb = ma;
    if (mt.na >= 1) ma = match_word(w, mt.a0);
    if (mt.na == 2) mb = match_word(w, mt.a0 + 1);
    if (ma.x == SFB_MATCH_SETS || mb.x == SFB_MATCH_SETS) return false;      // path sets: group_sets.cuh, or expanded first
    const int c0 = match_count(ma), c1 = match_count(mb);
    if (c0 + c1 > kR) return false;
    r.n = c0 + c1;
    r.c0 = c0;
    const uint2* buf = reinterpret_cast<const uint2*>(w.match_buf);
    const uint2* la = buf + (ma.x & 0x7fffffffu);
    const uint2* lb = buf + (mb.x & 0x7fffffffu);
#pragma unroll
    for (int i = 0; i < kR; ++i) {
        const bool second = i >= c0;
        const uint2 m = second ? mb : ma;
        uint2 t = m;
        if (i < r.n && (m.x & SFB_MATCH_MULTI)) t = second ? lb[i - c0] : la[i];
        r.code[i] = t.x;
        r.pid[i] = i < r.n ? (int)t.y : sentinel - i;
    }
    return true;
}

template <class T>
__device__ __forceinline__ T reg_pick(const T* v, int idx) {
    T out = v[0];
#pragma unroll
    for (int i = 1; i < kR; ++i)
        if (idx == i) out = v[i];
    return out;
}

struct PairMasks {
    unsigned in_other[2];   // bit i of mate k: its path id also occurs in the other mate
    unsigned dup[2];        // bit i: the path id occurs more than once in its own mate
    unsigned first[2];      // bit i: first tuple of its mate with that path id
};
__device__ __forceinline__ PairMasks pair_masks(const RegMate& a, const RegMate& b) {
    PairMasks m;
    m.in_other[0] = m.in_other[1] = m.dup[0] = m.dup[1] = 0;
    unsigned later0 = 0, later1 = 0;
#pragma unroll
    for (int i = 0; i < kR; ++i)
#pragma unroll
        for (int j = 0; j < kR; ++j) {
            const unsigned eq = a.pid[i] == b.pid[j];
            m.in_other[0] |= eq << i;
            m.in_other[1] |= eq << j;
        }
#pragma unroll
    for (int i = 0; i < kR; ++i)
#pragma unroll
        for (int j = i + 1; j < kR; ++j) {
            const unsigned e0 = a.pid[i] == a.pid[j], e1 = b.pid[i] == b.pid[j];
            m.dup[0] |= (e0 << i) | (e0 << j);
            m.dup[1] |= (e1 << i) | (e1 << j);
            later0 |= e0 << j;
            later1 |= e1 << j;
        }
    m.first[0] = ((1u << a.n) - 1u) & ~later0;
    m.first[1] = ((1u << b.n) - 1u) & ~later1;
    return m;
}

// strain words of every tuple (mask_words == 1), 0 for unused entries
__device__ __forceinline__ void strain_words(const sfb_graph& g, const RegMate& r, u64* sw) {
#pragma unroll
    for (int i = 0; i < kR; ++i) sw[i] = i < r.n ? g.path_smask[r.pid[i]] : 0ull;
}
__device__ __forceinline__ u64 or_all(const u64* sw) {
    u64 u = 0;
#pragma unroll
    for (int i = 0; i < kR; ++i) u |= sw[i];
    return u;
}

}  // namespace sfb
