// Shared device helpers for libsfb200 (sm_100a).  Internal; the public surface is include/sfb200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "sfb200.h"

namespace sfb {

typedef long long i64;
typedef unsigned long long u64;

int fail(int code, const char* fmt, ...);
int check_launch(const char* what);

#define SFB_REQUIRE(cond, msg)                                               \
    do {                                                                     \
        if (!(cond)) return ::sfb::fail(SFB_E_ARG, "%s: %s", __func__, msg); \
    } while (0)
#define SFB_TRY(x)                     \
    do {                               \
        int rc_ = (x);                 \
        if (rc_ != SFB_OK) return rc_; \
    } while (0)

constexpr int kWarp = 32;
constexpr int kSMs = 148;   // B200: persistent grids are sized in multiples of the SM count

static inline unsigned blocks_for(i64 n, int per) { return (unsigned)((n + per - 1) / per); }

// Grid of a persistent kernel: every SM filled with as many CTAs as are resident at once (one wave), never
// more CTAs than there is work.  The occupancy is queried once per kernel AND device: the cache is a table of
// atomics indexed by the device ordinal (a constant of the hardware, so racing writers store the same value).
struct OccCache {
    std::atomic<int> per_device[64];
    OccCache() { for (auto& v : per_device) v.store(0, std::memory_order_relaxed); }
};
template <class K>
static inline unsigned persistent_grid(K kernel, int block, i64 work_blocks, OccCache& cache) {
    int dev = 0;
    cudaGetDevice(&dev);
    int slots = dev >= 0 && dev < 64 ? cache.per_device[dev].load(std::memory_order_relaxed) : 0;
    if (slots == 0) {
        int per_sm = 0, sms = kSMs;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
        slots = per_sm * sms;
        if (dev >= 0 && dev < 64) cache.per_device[dev].store(slots, std::memory_order_relaxed);
    }
    const i64 g = work_blocks < (i64)slots ? work_blocks : (i64)slots;
    return (unsigned)(g < 1 ? 1 : g);
}

struct Dev {
    sfb_graph g;
    sfb_reads r;
    sfb_work w;
    sfb_accum acc;
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
// inclusive prefix sum over the warp
__device__ __forceinline__ unsigned warp_scan(unsigned v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= d) v += t;
    }
    return v;
}
// Reserve `need` items per lane on a global cursor with ONE atomic per warp; returns this lane's start.
// Must be called by all 32 lanes.
__device__ __forceinline__ u64 warp_claim(unsigned long long* cursor, unsigned need) {
    const unsigned incl = warp_scan(need);
    const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
    u64 base = 0;
    if (total) {
        if (lane_id() == 31) base = atomicAdd(cursor, (u64)total);
        base = __shfl_sync(0xffffffffu, base, 31);
    }
    return base + (incl - need);
}

// ---- block-local counters: shared-memory adds, one global add per counter per block ---------
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

// ---- match records (encoding in sfb200.h) -----------------------------------------------------
__device__ __forceinline__ uint2 match_word(const sfb_work& w, i64 a) {
    return reinterpret_cast<const uint2*>(w.aln_match)[a];
}
__device__ __forceinline__ int match_count(uint2 m) {          // (tuple form; path sets: match_count_at)
    if (m.x == SFB_MATCH_NONE) return 0;
    return (m.x & SFB_MATCH_MULTI) ? (int)m.y : 1;
}
// number of matches of alignment a, path sets included
__device__ __forceinline__ int match_count_at(const sfb_work& w, i64 a) {
    const uint2 m = match_word(w, a);
    if (m.x == SFB_MATCH_SETS) return __popcll(w.aln_sets[2 * a]) + __popcll(w.aln_sets[2 * a + 1]);
    return match_count(m);
}
// k-th match of an alignment as (code, path id)
__device__ __forceinline__ uint2 match_get(const sfb_work& w, uint2 m, int k) {
    if (!(m.x & SFB_MATCH_MULTI)) return m;
    return reinterpret_cast<const uint2*>(w.match_buf)[(size_t)(m.x & 0x7fffffffu) + k];
}

// Coverage of walk record `rec`: aligned bases / node length (json2csv.py:574), both packed by the decoder
// in one int so that the scatter kernels never touch the graph; or the explicit fp64 when the decoder had to
// merge mappings (:576-577).  The reference adds cov[i] to slot position+i whatever the orientation (Q3).
__device__ __forceinline__ double record_cov(const sfb_reads& r, i64 rec) {
    const int v = r.walk_alen[rec];
    if (v < 0) return r.exc_cov[-(i64)v - 1];
    return (double)(v & 0xffff) / (double)(v >> 16);        // aligned bases / node length, IEEE division
}

// ---- order-independent accumulation: 96-bit fixed point ------------------------------------------------------
// A term v >= 0 (v < 2^32) is the integer floor(v * 2^64), split into three 32-bit limbs x2:x1:x0 (x2 = integer
// part, x1 = multiples of 2^-32, x0 = multiples of 2^-64).  Integer sums are exact, so the result does not depend
// on the order of the additions, on the sharding of the reads or on the number of GPUs.  In global memory an
// accumulator is two 64-bit integer planes: plane 0 (at x) holds x2:x1 in units of 2^-32, plane 1 (at x +
// det_stride) holds the x0 limbs in units of 2^-64; sfb_accum_collapse turns the pair into the fp64 value.
struct Fx {
    uint32_t x0, x1, x2;
};
__device__ __forceinline__ Fx fx_split(double v) {
    const double t = v * 4294967296.0;                 // exact (power of two)
    const double hi = floor(t);
    const u64 H = __double2ull_rz(hi);
    Fx f;
    f.x0 = __double2uint_rz((t - hi) * 4294967296.0);  // t - hi is exact and in [0, 1)
    f.x1 = (uint32_t)H;
    f.x2 = (uint32_t)(H >> 32);
    return f;
}
__device__ __forceinline__ void fx_red(double* p, i64 det_stride, Fx f) {
    const u64 H = ((u64)f.x2 << 32) | f.x1;
    if (H) atomicAdd(reinterpret_cast<u64*>(p), H);
    if (f.x0) atomicAdd(reinterpret_cast<u64*>(p + det_stride), (u64)f.x0);
}
// fp64 accumulation: plain RED (det_stride == 0: the order of the additions, hence the last bits, vary from run to
// run), or the two exact integer planes above.  v >= 0.
__device__ __forceinline__ void acc_add(double* p, i64 det_stride, double v) {
    if (det_stride == 0) {
        atomicAdd(p, v);
        return;
    }
    fx_red(p, det_stride, fx_split(v));
}

// slot -> path id from the 64-slot block table: one 8-byte load; path_off is only walked for blocks in which several
// paths start (paths shorter than 64 nodes)
__device__ __forceinline__ int path_of_block(const sfb_graph& g, int2 e, int s) {
    int p = e.x;
    if (e.y >= 0) return p + (s >= e.y);
    while (g.path_off[p + 1] <= s) ++p;
    return p;
}
__device__ __forceinline__ int path_of_slot(const sfb_graph& g, int s) {
    return path_of_block(g, reinterpret_cast<const int2*>(g.blk_path)[s >> 6], s);
}

// Path sets (SFB_MATCH_SETS, sfb200.h): the tuple of the lowest set bit of `rest`, a subset of the forward (reverse) set of
// an alignment whose walk starts (ends) on `node`.  Occurrences are ordered by slot, hence by path: the rank of the bit
// inside node_mask is the rank of the occurrence.
__device__ __forceinline__ uint2 sets_tuple(const sfb_graph& g, int node, int p0, u64 rest, bool rev) {
    const u64 bit = rest & (~rest + 1);
    const int k = __popcll(g.node_mask[node] & (bit - 1));
    const int slot = reinterpret_cast<const int2*>(g.occ_pair)[g.occ_off[node] + k].x;
    return make_uint2((uint32_t)slot | (rev ? SFB_MATCH_REV : 0u), (uint32_t)(p0 + __ffsll((long long)bit) - 1));
}

// A mate of a read group: its retained alignments [a0, a0+na) (json2csv.py:686-739).
struct Mate {
    i64 a0;
    int na;
    int seq_len;
};
__device__ __forceinline__ void mate_load(const sfb_reads& r, i64 g, int m, Mate& out) {
    out.a0 = r.grp_aln_off[2 * g + m];
    out.na = (int)(r.grp_aln_off[2 * g + m + 1] - out.a0);
    out.seq_len = r.mate_seq_len[2 * g + m];
}

// ---- views of a group's match tuples ("colored_match", json2csv.py:851-866) ---------------------
// Tuple i of mate k = (code, path id, local alignment index), in the reference's order.
constexpr int kCap = 32;   // tuples per mate held in thread-local memory by the fast view

struct LocalView {
    int n[2];
    int pid_[2][kCap];
    uint32_t code_[2][kCap];
    unsigned char aloc_[2][kCap];
    __device__ __forceinline__ int nt(int k) const { return n[k]; }
    __device__ __forceinline__ int pid(int k, int i) const { return pid_[k][i]; }
    __device__ __forceinline__ uint32_t code(int k, int i) const { return code_[k][i]; }
    __device__ __forceinline__ int aloc(int k, int i) const { return aloc_[k][i]; }
    // false when the group does not fit (then use GlobalView)
    __device__ __forceinline__ bool load(const sfb_work& w, const Mate* mt) {
        for (int k = 0; k < 2; ++k) {
            int c = 0;
            if (mt[k].na > 255) return false;
            for (int j = 0; j < mt[k].na; ++j) {
                const uint2 m = match_word(w, mt[k].a0 + j);
                const int cnt = match_count(m);
                if (c + cnt > kCap) return false;
                for (int e = 0; e < cnt; ++e, ++c) {
                    const uint2 t = match_get(w, m, e);
                    code_[k][c] = t.x;
                    pid_[k][c] = (int)t.y;
                    aloc_[k][c] = (unsigned char)j;
                }
            }
            n[k] = c;
        }
        return true;
    }
};

// Same interface straight from global memory, for groups with very many matches (any size).
struct GlobalView {
    const sfb_work* w;
    Mate mt[2];
    int n[2];
    __device__ __forceinline__ void init(const sfb_work& w_, const Mate* m) {
        w = &w_;
        for (int k = 0; k < 2; ++k) {
            mt[k] = m[k];
            int c = 0;
            for (int j = 0; j < m[k].na; ++j) c += match_count(match_word(w_, m[k].a0 + j));
            n[k] = c;
        }
    }
    __device__ __forceinline__ uint2 at(int k, int i, int& j) const {
        for (j = 0; j < mt[k].na; ++j) {
            const uint2 m = match_word(*w, mt[k].a0 + j);
            const int c = match_count(m);
            if (i < c) return match_get(*w, m, i);
            i -= c;
        }
        return make_uint2(0, 0);
    }
    __device__ __forceinline__ int nt(int k) const { return n[k]; }
    __device__ __forceinline__ int pid(int k, int i) const { int j; return (int)at(k, i, j).y; }
    __device__ __forceinline__ uint32_t code(int k, int i) const { int j; return at(k, i, j).x; }
    __device__ __forceinline__ int aloc(int k, int i) const { int j; at(k, i, j); return j; }
};

// list_path_id.count(pid) (json2csv.py:932)
template <class V>
__device__ __forceinline__ int count_pid(const V& v, int k, int p) {
    int n = 0;
    for (int i = 0; i < v.nt(k); ++i) n += v.pid(k, i) == p;
    return n;
}
// tuple i is the first of mate k carrying its path id
template <class V>
__device__ __forceinline__ bool first_of_pid(const V& v, int k, int i) {
    const int p = v.pid(k, i);
    for (int j = 0; j < i; ++j)
        if (v.pid(k, j) == p) return false;
    return true;
}
// strains common to both mates' candidate paths, one 64-bit word of the bitmasks (json2csv.py:964-972)
template <class V>
__device__ __forceinline__ u64 common_strains(const sfb_graph& g, const V& v, int word) {
    u64 u[2] = {0, 0};
    for (int k = 0; k < 2; ++k)
        for (int i = 0; i < v.nt(k); ++i) u[k] |= g.path_smask[(size_t)v.pid(k, i) * g.mask_words + word];
    return u[0] & u[1];
}
// how many times path p appears in the strain-reduced list (once per common strain it carries, Q4)
template <class V>
__device__ __forceinline__ int strain_times(const sfb_graph& g, const V& v, int p) {
    int m = 0;
    for (int word = 0; word < (int)g.mask_words; ++word)
        m += __popcll(g.path_smask[(size_t)p * g.mask_words + word] & common_strains(g, v, word));
    return m;
}

// pass-2 work item: add weight * cov[i] to mult_abund[slot + i] for every record of alignment `aln`
struct __align__(16) App {
    int32_t aln;
    uint32_t code;
    double weight;
};

}  // namespace sfb
