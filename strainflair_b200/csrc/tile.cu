// Tile kernels: the abundance pass for read groups that live inside one tile of the graph (sfb_tiles in sfb200.h).
//
// A thread block takes a work item (tile, range of read groups in tile order), stages the tile in shared memory --
// path nodes, node -> occurrence index, slot -> path, per-path constants -- together with zeroed accumulator bins,
// and then runs, one thread per read group, everything the reference does for the group:
//   k_tile_pass1   get_matching_path (json2csv.py:341-363) on both orientations of every alignment, the pass-1
//                  decision tree (:828-1053), fill_abund_dist (:321-339) including its node loop
//   k_tile_pass2   for the deferred groups: the matches again (as the reference recomputes them, :1103-1114), the
//                  pass-2 candidate selection and ratios (:1123-1274), the node loops (:1157-1158)
// Compares run against shared memory instead of index gathers, the fp64 additions go to shared-memory bins in 96-bit
// fixed point (three 32-bit limbs per accumulator, native ATOMS.ADD with exact carries: integer sums do not depend
// on the order of the additions) and are flushed once per work item with 64-bit integer REDs; histogram updates are
// aggregated per block in a small hash table keyed by (strain, five bins).  The decision logic is group_logic.cuh,
// shared with the global-memory kernels.
#include <atomic>

#include "common.cuh"
#include "group_logic.cuh"

namespace sfb {

constexpr int kTT = 256;              // threads per block
constexpr int kHash = 256;            // entries of the histogram aggregation table
constexpr u64 kHashEmpty = ~0ull;
constexpr unsigned kNoNode = 0xFFFEu; // local id of a walk node outside the tile (never equal to a path node)
constexpr unsigned kSentinel = 0xFFFFu;

// ---- shared-memory layout, sized by the largest tile (same arithmetic on host and device) -----------------------
struct TileLayout {
    unsigned hkey, pseq, pU, tup, bins, plimb, pclu, pstr0, pnstr, hcnt, sc, misc, poff, node, occ, occ_off, spath, total;
};
__host__ __device__ inline TileLayout tile_layout(int ms, int mn, int mp, int K) {
    TileLayout L;
    unsigned o = 0;
    auto take = [&](unsigned bytes) { const unsigned at = o; o += (bytes + 15u) & ~15u; return at; };
    L.hkey = take(8u * kHash);
    L.pseq = take(8u * mp);
    L.pU = take(8u * mp);
    L.tup = take(4u * 2 * K * kTT);
    L.bins = take(12u * ms);
    L.plimb = take(28u * mp);
    L.pclu = take(4u * mp);
    L.pstr0 = take(4u * mp);
    L.pnstr = take(4u * mp);
    L.hcnt = take(4u * kHash);
    L.sc = take(sizeof(BlockCounters));
    L.misc = take(16);
    L.poff = take(2u * (mp + 1));
    L.node = take(2u * (ms + 2));
    L.occ = take(2u * ms);
    L.occ_off = take(2u * (mn + 1));
    L.spath = take(ms);
    L.total = o;
    return L;
}

struct TileS {
    u64* hkey;
    long long* pseq;
    long long* pU;
    uint32_t* tup;       // [2][K][kTT], this thread's column
    uint32_t* bins;      // [ns][3] limbs x0, x1, x2
    uint32_t* plimb;     // [np][7]: count / flag, three limbs A, three limbs B
    int* pclu;
    int* pstr0;
    int* pnstr;
    uint32_t* hcnt;
    BlockCounters* sc;
    int* misc;
    uint16_t* poff;
    uint16_t* node;
    uint16_t* occ;
    uint16_t* occ_off;
    uint8_t* spath;
    int p0, np, n0, nn, s0, ns;
};
__device__ __forceinline__ TileS tile_carve(unsigned char* smem, const sfb_tiles& t) {
    const TileLayout L = tile_layout((int)t.max_slots, (int)t.max_nodes, (int)t.max_paths, (int)t.tuple_cap);
    TileS T;
    T.hkey = reinterpret_cast<u64*>(smem + L.hkey);
    T.pseq = reinterpret_cast<long long*>(smem + L.pseq);
    T.pU = reinterpret_cast<long long*>(smem + L.pU);
    T.tup = reinterpret_cast<uint32_t*>(smem + L.tup) + threadIdx.x;
    T.bins = reinterpret_cast<uint32_t*>(smem + L.bins);
    T.plimb = reinterpret_cast<uint32_t*>(smem + L.plimb);
    T.pclu = reinterpret_cast<int*>(smem + L.pclu);
    T.pstr0 = reinterpret_cast<int*>(smem + L.pstr0);
    T.pnstr = reinterpret_cast<int*>(smem + L.pnstr);
    T.hcnt = reinterpret_cast<uint32_t*>(smem + L.hcnt);
    T.sc = reinterpret_cast<BlockCounters*>(smem + L.sc);
    T.misc = reinterpret_cast<int*>(smem + L.misc);
    T.poff = reinterpret_cast<uint16_t*>(smem + L.poff);
    T.node = reinterpret_cast<uint16_t*>(smem + L.node);
    T.occ = reinterpret_cast<uint16_t*>(smem + L.occ);
    T.occ_off = reinterpret_cast<uint16_t*>(smem + L.occ_off);
    T.spath = smem + L.spath;
    T.p0 = T.np = T.n0 = T.nn = T.s0 = T.ns = 0;
    return T;
}

struct TileStats {
    int cand, cmp, tuples, rec, app, deferred;
};

// ---- fixed-point bins in shared memory ---------------------------------------------------------------------------
// b[0..2] += (x0, x1, x2) as one 96-bit integer: every limb add is a native 32-bit ATOMS.ADD that returns the old
// value, a wrap of a limb is carried into the next one by the thread that caused it -- exactly once per wrap, so
// the 96-bit total is exact whatever the interleaving of the threads.
__device__ __forceinline__ void fx_smem_add(uint32_t* b, Fx f) {
    uint32_t carry = 0;
    if (f.x0) {
        const uint32_t old = atomicAdd(b, f.x0);
        carry = (uint32_t)(old + f.x0 < old);
    }
    uint32_t t1 = f.x1 + carry;
    uint32_t c1 = (uint32_t)(t1 < carry);          // x1 = 2^32 - 1 and a carry: t1 wraps to 0
    if (t1) {
        const uint32_t old = atomicAdd(b + 1, t1);
        c1 |= (uint32_t)(old + t1 < old);
    }
    const uint32_t t2 = f.x2 + c1;
    if (t2) atomicAdd(b + 2, t2);
}
// one accumulator out to the global planes (or to the fp64 array in the plain mode)
__device__ __forceinline__ void fx_flush(double* p, i64 det_stride, const uint32_t* b) {
    const uint32_t x0 = b[0], x1 = b[1], x2 = b[2];
    if ((x0 | x1 | x2) == 0) return;
    if (det_stride) {
        Fx f;
        f.x0 = x0; f.x1 = x1; f.x2 = x2;
        fx_red(p, det_stride, f);
    } else {
        atomicAdd(p, (double)x2 + (double)x1 * 0x1p-32 + (double)x0 * 0x1p-64);
    }
}
// cov[i] * weight into the bins of slot (local); full-coverage records (aligned == node length) with weight 1 are
// the integer 1
__device__ __forceinline__ void tile_add_records(const TileS& T, const sfb_reads& r, i64 w0, int L, int slot, bool unit,
                                                 double weight) {
    uint32_t* b = T.bins + 3 * slot;
    for (int i = 0; i < L; ++i, b += 3) {
        const int v = r.walk_alen[w0 + i];
        if (unit && v >= 0 && (v & 0xffff) == (v >> 16)) {
            atomicAdd(b + 2, 1u);
            continue;
        }
        const double cov = v < 0 ? r.exc_cov[-(i64)v - 1] : (double)(v & 0xffff) / (double)(v >> 16);
        fx_smem_add(b, fx_split(unit ? cov : cov * weight));
    }
}

// ---- staging / flushing of a tile -------------------------------------------------------------------------------
template <int PASS>
__device__ __forceinline__ void tile_stage(TileS& T, const Dev& d, const sfb_tiles& t, int tile, const long long* U) {
    const int4 desc = reinterpret_cast<const int4*>(t.tile_desc)[tile];
    T.p0 = desc.x;
    T.np = desc.y - desc.x;
    T.n0 = desc.z;
    T.nn = desc.w - desc.z;
    T.s0 = d.g.path_off[T.p0];
    T.ns = d.g.path_off[desc.y] - T.s0;
    const int tid = threadIdx.x;
    for (int i = tid; i < T.ns; i += kTT) {
        const int v = d.g.slot_node[T.s0 + i];
        T.node[i] = v < 0 ? (uint16_t)kSentinel : (uint16_t)(v - T.n0);
    }
    if (tid < 2) T.node[T.ns + tid] = (uint16_t)kSentinel;
    const int obase = d.g.occ_off[T.n0];
    for (int i = tid; i <= T.nn; i += kTT) T.occ_off[i] = (uint16_t)(d.g.occ_off[T.n0 + i] - obase);
    const int nocc = d.g.occ_off[T.n0 + T.nn] - obase;
    const int2* occ = reinterpret_cast<const int2*>(d.g.occ_pair);
    for (int i = tid; i < nocc; i += kTT) T.occ[i] = (uint16_t)(occ[obase + i].x - T.s0);
    for (int p = tid; p <= T.np; p += kTT) T.poff[p] = (uint16_t)(d.g.path_off[T.p0 + p] - T.s0);
    for (int p = tid; p < T.np; p += kTT) {
        T.pseq[p] = d.g.path_seq_len[T.p0 + p];
        T.pclu[p] = d.g.path_cluster[T.p0 + p];
        T.pstr0[p] = d.g.path_strain0[T.p0 + p];
        T.pnstr[p] = d.g.path_nstrain[T.p0 + p];
        if (PASS == 2) T.pU[p] = U[T.p0 + p];
    }
    for (int i = tid; i < 3 * T.ns; i += kTT) T.bins[i] = 0u;
    for (int i = tid; i < 7 * T.np; i += kTT) T.plimb[i] = 0u;
    for (int i = tid; i < kHash; i += kTT) {
        T.hkey[i] = kHashEmpty;
        T.hcnt[i] = 0u;
    }
    __syncthreads();
    // slot -> local path: binary search over the staged path offsets
    for (int i = tid; i < T.ns; i += kTT) {
        int lo = 0, hi = T.np;               // largest p with poff[p] <= i
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if ((int)T.poff[mid] <= i) lo = mid; else hi = mid;
        }
        T.spath[i] = (uint8_t)lo;
    }
    __syncthreads();
}

__device__ __forceinline__ void hist_out(const Dev& d, BlockCounters& sc, u64 key, uint32_t cnt) {
    const size_t s = (size_t)(key >> 40);
    SFB_COUNT(sc, SFB_C_ST_HIST, (int)cnt);
    for (int k = 0; k < SFB_N_HIST; ++k) {
        const int b = (int)((key >> (8 * k)) & 0xffu);
        if (b >= SFB_N_BINS) { SFB_COUNT(sc, SFB_C_BAD_BIN, (int)cnt); continue; }
        atomicAdd(&d.acc.hist[((size_t)k * d.g.n_strains + s) * SFB_N_BINS + b], (u64)cnt);
    }
}

template <int PASS>
__device__ __forceinline__ void tile_flush(const TileS& T, const Dev& d, BlockCounters& sc) {
    const int tid = threadIdx.x;
    double* slots = PASS == 1 ? d.acc.uniq_abund : d.acc.mult_abund;
    for (int i = tid; i < T.ns; i += kTT) fx_flush(slots + T.s0 + i, d.acc.det_stride, T.bins + 3 * i);
    for (int p = tid; p < T.np; p += kTT) {
        const uint32_t* l = T.plimb + 7 * p;
        const int pid = T.p0 + p;
        if (PASS == 1) {
            if (l[0]) atomicAdd((u64*)&d.acc.uniq_reads[pid], (u64)l[0]);
            fx_flush(&d.acc.uniq_norm[pid], d.acc.det_stride, l + 1);
        } else {
            if (l[0]) d.acc.mult_hits[pid] = 1u;
            fx_flush(&d.acc.mult_reads[pid], d.acc.det_stride, l + 1);
            fx_flush(&d.acc.mult_norm[pid], d.acc.det_stride, l + 4);
        }
    }
    for (int i = tid; i < kHash; i += kTT)
        if (T.hkey[i] != kHashEmpty) hist_out(d, sc, T.hkey[i], T.hcnt[i]);
}

// ---- matching against the resident tile -------------------------------------------------------------------------
__device__ __forceinline__ unsigned tile_local(const TileS& T, int node) {
    const unsigned x = (unsigned)(node - T.n0);
    return x < (unsigned)T.nn ? x : kNoNode;
}
// get_matching_path (json2csv.py:341-363) for the walk and, if longer than one node, the reversed walk: f(slot, rev)
// for every match, in the reference's order (forward by ascending slot, then reverse)
template <class F>
__device__ __forceinline__ void tile_enum(const TileS& T, const sfb_reads& r, i64 w0, int L, TileStats& st, F&& f) {
    if (L <= 0) return;
    const unsigned first = tile_local(T, r.walk_node[w0]);
    if (L == 1) {
        if (first == kNoNode) return;
        for (int c = T.occ_off[first]; c < (int)T.occ_off[first + 1]; ++c) {
            st.cand += 1;
            f((int)T.occ[c], false);
        }
        return;
    }
    const unsigned last = tile_local(T, r.walk_node[w0 + L - 1]);
    const unsigned second = tile_local(T, r.walk_node[w0 + 1]), penult = tile_local(T, r.walk_node[w0 + L - 2]);
    if (first != kNoNode)
        for (int c = T.occ_off[first]; c < (int)T.occ_off[first + 1]; ++c) {
            const int s = T.occ[c];
            st.cand += 1;
            st.cmp += 1;
            if (T.node[s + 1] != second) continue;
            bool ok = true;
            for (int i = 2; i < L && ok; ++i) {
                st.cmp += 1;
                ok = T.node[s + i] == tile_local(T, r.walk_node[w0 + i]);     // a sentinel ends the compare
            }
            if (ok) f(s, false);
        }
    if (last != kNoNode)
        for (int c = T.occ_off[last]; c < (int)T.occ_off[last + 1]; ++c) {
            const int s = T.occ[c];
            st.cand += 1;
            st.cmp += 1;
            if (T.node[s + 1] != penult) continue;
            bool ok = true;
            for (int i = 2; i < L && ok; ++i) {
                st.cmp += 1;
                ok = T.node[s + i] == tile_local(T, r.walk_node[w0 + L - 1 - i]);
            }
            if (ok) f(s, true);
        }
}

// match tuple in shared memory: rev (1) | local slot (15) | local path (8) | alignment inside the mate (8)
__device__ __forceinline__ uint32_t tup_pack(int slot, int pidx, int aloc, bool rev) {
    return ((uint32_t)rev << 31) | ((uint32_t)slot << 16) | ((uint32_t)pidx << 8) | (uint32_t)aloc;
}
__device__ __forceinline__ uint32_t tile_code(const TileS& T, int slot, bool rev) {
    return (uint32_t)(T.s0 + slot) | (rev ? SFB_MATCH_REV : 0u);
}

template <int K>
struct TileView {
    const TileS* T;
    int n[2];
    __device__ __forceinline__ uint32_t word(int k, int i) const { return T->tup[(k * K + i) * kTT]; }
    __device__ __forceinline__ int nt(int k) const { return n[k]; }
    __device__ __forceinline__ int pid(int k, int i) const { return T->p0 + (int)((word(k, i) >> 8) & 0xffu); }
    __device__ __forceinline__ uint32_t code(int k, int i) const {
        const uint32_t w = word(k, i);
        return tile_code(*T, (int)((w >> 16) & 0x7fffu), (w >> 31) != 0);
    }
    __device__ __forceinline__ int aloc(int k, int i) const { return (int)(word(k, i) & 0xffu); }
};

// tuples of both mates of an aligned pair into shared memory; false when a mate has more than K (or more than 255
// alignments): the pair goes to the global-memory kernels
template <int K>
__device__ __forceinline__ bool tile_match_pair(const TileS& T, const sfb_reads& r, const Mate* mt, TileView<K>& v,
                                                TileStats& st) {
    bool fits = true;
    for (int k = 0; k < 2; ++k) {
        int n = 0;
        if (mt[k].na > 255) fits = false;
        for (int j = 0; j < mt[k].na && fits; ++j) {
            const i64 w0 = r.aln_walk_off[mt[k].a0 + j];
            const int L = (int)(r.aln_walk_off[mt[k].a0 + j + 1] - w0);
            tile_enum(T, r, w0, L, st, [&](int s, bool rev) {
                if (n < K) T.tup[(k * K + n) * kTT] = tup_pack(s, T.spath[s], j, rev);
                ++n;
            });
            if (n > K) fits = false;
        }
        v.n[k] = n;
        st.tuples += n;
    }
    return fits;
}

// ---- effects (the Ops of group_logic.cuh) on the resident tile ---------------------------------------------------
struct TileOps {
    TileS* T;
    TileStats* st;
    __device__ __forceinline__ void hist(const Dev& d, BlockCounters& sc, int pid, i64 a) {
        const int p = pid - T->p0;
        if (T->pnstr[p] != 1) return;                        // json2csv.py:333 / :1246
        const uint8_t* b = d.r.aln_bins + a * SFB_N_HIST;
        const u64 key = ((u64)(uint32_t)T->pstr0[p] << 40) | (u64)b[0] | ((u64)b[1] << 8) | ((u64)b[2] << 16) |
                        ((u64)b[3] << 24) | ((u64)b[4] << 32);
        unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 56) & (kHash - 1);
        for (int probe = 0; probe < kHash; ++probe, h = (h + 1) & (kHash - 1)) {
            const u64 old = atomicCAS(&T->hkey[h], kHashEmpty, key);
            if (old == kHashEmpty || old == key) {
                atomicAdd(&T->hcnt[h], 1u);
                return;
            }
        }
        hist_out(d, sc, key, 1u);                            // table full: straight to the global histograms
    }
    // fill_abund_dist (json2csv.py:321-339), node loop included
    __device__ __forceinline__ void assign_unique(const Dev& d, BlockCounters& sc, i64 a, uint32_t code, int pid, int seq_len) {
        d.w.aln_assign[a] = (int32_t)code;
        const int p = pid - T->p0;
        atomicAdd(&T->plimb[7 * p], 1u);
        fx_smem_add(&T->plimb[7 * p + 1], fx_split((double)seq_len / (double)T->pseq[p]));
        hist(d, sc, pid, a);
        const i64 w0 = d.r.aln_walk_off[a];
        const int L = (int)(d.r.aln_walk_off[a + 1] - w0);
        tile_add_records(*T, d.r, w0, L, (int)(code & SFB_SLOT_MASK) - T->s0, true, 1.0);
        st->rec += L;
    }
    // cluster of the first path through the walk's first node (json2csv.py:875,901,1036); -1 = None, -2 = no path
    __device__ __forceinline__ int first_cluster(const Dev& d, i64 a) {
        const unsigned n = tile_local(*T, d.r.walk_node[d.r.aln_walk_off[a]]);
        if (n == kNoNode || T->occ_off[n] == T->occ_off[n + 1]) return -2;
        return T->pclu[T->spath[T->occ[T->occ_off[n]]]];
    }
    __device__ __forceinline__ void book(const Dev& d, BlockCounters& sc, const Mate& mt, int cluster) {
        if (cluster < 0) { SFB_COUNT(sc, SFB_C_NO_CLUSTER, 1); return; }
        d.w.aln_assign[mt.a0] = -(cluster + 2);
    }
    __device__ __forceinline__ long long uniq(int pid) const { return T->pU[pid - T->p0]; }
    // json2csv.py:1155-1158 (and :1207-1210, :1265-1268)
    __device__ __forceinline__ void apply(const Dev& d, i64 a, uint32_t code, int pid, int seq_len, double ratio, int times) {
        const int p = pid - T->p0;
        const double weight = ratio * (double)times;
        T->plimb[7 * p] = 1u;                                // "touched" flag
        fx_smem_add(&T->plimb[7 * p + 1], fx_split(weight));
        fx_smem_add(&T->plimb[7 * p + 4], fx_split((ratio * (double)seq_len / (double)T->pseq[p]) * (double)times));
        const i64 w0 = d.r.aln_walk_off[a];
        const int L = (int)(d.r.aln_walk_off[a + 1] - w0);
        tile_add_records(*T, d.r, w0, L, (int)(code & SFB_SLOT_MASK) - T->s0, false, weight);
        st->rec += L;
        st->app += 1;
    }
};

// An aligned pair that does not fit the tuple buffers: its match lists in the format of sfb_match, the group on the
// list of the generic classify kernel (which also takes care of its node loops and of its deferral).
__device__ __noinline__ void tile_export(const TileS& T, const Dev& d, BlockCounters& sc, i64 g, const Mate* mt,
                                         TileStats& st) {
    uint2* aln_match = reinterpret_cast<uint2*>(d.w.aln_match);
    uint2* match_buf = reinterpret_cast<uint2*>(d.w.match_buf);
    for (int k = 0; k < 2; ++k)
        for (int j = 0; j < mt[k].na; ++j) {
            const i64 a = mt[k].a0 + j;
            const i64 w0 = d.r.aln_walk_off[a];
            const int L = (int)(d.r.aln_walk_off[a + 1] - w0);
            int n = 0;
            uint2 only = make_uint2(SFB_MATCH_NONE, 0u);
            tile_enum(T, d.r, w0, L, st, [&](int s, bool rev) {
                if (n++ == 0) only = make_uint2(tile_code(T, s, rev), (uint32_t)(T.p0 + T.spath[s]));
            });
            if (n <= 1) {
                aln_match[a] = only;
                continue;
            }
            const u64 pos = atomicAdd(d.w.cursor, (u64)n);
            if (pos + n > (u64)d.w.match_cap) {
                aln_match[a] = make_uint2(SFB_MATCH_NONE, 0u);
                SFB_COUNT(sc, SFB_C_MATCH_OVERFLOW, n);
                continue;
            }
            int q = 0;
            tile_enum(T, d.r, w0, L, st, [&](int s, bool rev) {
                match_buf[pos + q++] = make_uint2(tile_code(T, s, rev), (uint32_t)(T.p0 + T.spath[s]));
            });
            aln_match[a] = make_uint2(SFB_MATCH_MULTI | (uint32_t)pos, (uint32_t)n);
        }
    d.w.slow[atomicAdd(d.w.cursor + 3, 1ull)] = (int32_t)g;
}

// ---- pass 1, one read group (json2csv.py:828-1053) --------------------------------------------------------------
template <int K>
__device__ __forceinline__ void tile_group1(TileS& T, const Dev& d, BlockCounters& sc, i64 g, TileStats& st) {
    const int nm = d.r.grp_nmates[g];
    Mate mt[2];
    mate_load(d.r, g, 0, mt[0]);
    mate_load(d.r, g, 1, mt[1]);
    for (int k = 0; k < 2; ++k)
        for (int j = 0; j < mt[k].na; ++j) d.w.aln_assign[mt[k].a0 + j] = -1;
    int code[2] = {SFB_ABSENT, SFB_ABSENT};
    bool defer = false;
    SFB_COUNT(sc, SFB_C_NB_READS, nm);
    int single = nm == 1 ? 0 : -1;           // mate handled single-end style, -1 = none
    int known = -1;                          // its number of matches when already counted
    TileOps ops{&T, &st};
    if (nm == 2) {
        SFB_COUNT(sc, SFB_C_PAIRS, 1);
        if (mt[0].na == 0 && mt[1].na == 0) {
            SFB_COUNT(sc, SFB_C_FILTERED, 2);
            code[0] = code[1] = SFB_FILTERED;
        } else if (mt[0].na == 0 || mt[1].na == 0) {
            SFB_COUNT(sc, SFB_C_FILTERED, 1);
            SFB_COUNT(sc, SFB_C_ONE_NOALIGN, 1);
            const int gone = mt[0].na == 0 ? 0 : 1;
            code[gone] = SFB_FILTERED;
            single = 1 - gone;
        } else {
            TileView<K> v;
            v.T = &T;
            if (!tile_match_pair<K>(T, d.r, mt, v, st)) {
                tile_export(T, d, sc, g, mt, st);
                return;
            }
            if (v.n[0] == 0 && v.n[1] == 0) {
                classify_both_unassigned(d, ops, sc, mt, code);
            } else if (v.n[0] == 0 || v.n[1] == 0) {
                // one mate without colored path (json2csv.py:895-916)
                SFB_COUNT(sc, SFB_C_UNASSIGNED, 1);
                SFB_COUNT(sc, SFB_C_ONE_NOCP, 1);
                const int gone = v.n[0] == 0 ? 0 : 1;
                code[gone] = SFB_UNASSIGNED;
                bool bad = false;
                for (int j = 0; j < mt[gone].na; ++j) bad |= ops.first_cluster(d, mt[gone].a0 + j) == -2;
                if (bad) {
                    SFB_COUNT(sc, SFB_C_NO_PATH, 1);
                } else if (mt[gone].na == 1) {
                    ops.book(d, sc, mt[gone], ops.first_cluster(d, mt[gone].a0));
                    code[gone] = SFB_UNASSIGNED_BOOK;
                }
                single = 1 - gone;
                known = v.n[single];
            } else {
                classify_pair(d, ops, sc, v, mt, code, defer);
            }
        }
    }
    if (single >= 0) {
        // single-end style (json2csv.py:1003-1050)
        const Mate& m = mt[single];
        SFB_COUNT(sc, SFB_C_SINGLES, 1);
        if (m.na == 0) {
            SFB_COUNT(sc, SFB_C_FILTERED, 1);
            code[single] = SFB_FILTERED;
        } else if (m.na > 1) {
            SFB_COUNT(sc, SFB_C_SECOND_PASS, 1);
            defer = true;
            code[single] = SFB_DEFER;
        } else {
            int n = known;
            if (n < 0) {
                n = 0;
                const i64 w0 = d.r.aln_walk_off[m.a0];
                tile_enum(T, d.r, w0, (int)(d.r.aln_walk_off[m.a0 + 1] - w0), st, [&](int, bool) { ++n; });
                st.tuples += n;
            }
            if (n > 1) {
                SFB_COUNT(sc, SFB_C_SECOND_PASS, 1);
                defer = true;
                code[single] = SFB_DEFER;
            } else if (n == 0) {
                SFB_COUNT(sc, SFB_C_UNASSIGNED, 1);
                const int c = ops.first_cluster(d, m.a0);
                if (c == -2) SFB_COUNT(sc, SFB_C_NO_PATH, 1); else ops.book(d, sc, m, c);
                code[single] = SFB_UNASSIGNED_BOOK;
            } else {
                SFB_COUNT(sc, SFB_C_ASSIGNED_U, 1);      // counted, never accumulated (json2csv.py:1047-1053, Q1)
                SFB_COUNT(sc, SFB_C_SE, 1);
                code[single] = SFB_SE_COUNTED;
            }
        }
    }
    d.w.grp_code[g] = (uint8_t)(code[0] | (code[1] << 4));
    st.deferred += defer;
}

// ---- pass 2, one deferred read group (json2csv.py:1085-1274) ----------------------------------------------------
template <int K>
__device__ __forceinline__ void tile_group2(TileS& T, const Dev& d, BlockCounters& sc, i64 g, TileStats& st) {
    const int c1 = d.w.grp_code[g];
    if ((c1 & 0xf) != SFB_DEFER && (c1 >> 4) != SFB_DEFER) return;
    const int nm = d.r.grp_nmates[g];
    Mate mt[2];
    mate_load(d.r, g, 0, mt[0]);
    mate_load(d.r, g, 1, mt[1]);
    int code[2] = {SFB_P2_NONE, SFB_P2_NONE};
    int single = nm == 1 ? 0 : -1;
    TileOps ops{&T, &st};
    if (nm == 2) {
        if (mt[0].na == 0 || mt[1].na == 0) {
            single = mt[0].na == 0 ? 1 : 0;                                  // json2csv.py:1088-1093
        } else {
            TileView<K> v;
            v.T = &T;
            const int seen = st.tuples;
            const bool fits = tile_match_pair<K>(T, d.r, mt, v, st);
            st.tuples = seen;                                                // counted in pass 1
            if (!fits) return;                                               // exported in pass 1: sfb_pass2 has it
            if (v.n[0] == 0 || v.n[1] == 0) {
                single = v.n[0] == 0 ? 1 : 0;                                // :1117-1122
            } else {
                Plan pl;
                pl.mode[0] = pl.mode[1] = kNone;
                pl.napp[0] = pl.napp[1] = 0;
                p2_decide_pair(d, ops, sc, v, pl, code);
                p2_emit(d, ops, sc, v, mt, pl);
            }
        }
    }
    if (single >= 0) {
        // every alignment separately (json2csv.py:1230-1268): the matches are enumerated twice instead of stored
        const Mate& m = mt[single];
        int n_all = 0;
        for (int j = 0; j < m.na; ++j) {
            const i64 a = m.a0 + j;
            const i64 w0 = d.r.aln_walk_off[a];
            const int L = (int)(d.r.aln_walk_off[a + 1] - w0);
            long long total = 0;
            int cnt = 0;
            tile_enum(T, d.r, w0, L, st, [&](int s, bool) {
                const int p = T.spath[s];
                total += T.pU[p];
                ++cnt;
                ops.hist(d, sc, T.p0 + p, a);                                // :1246-1252
            });
            if (cnt == 0) continue;
            n_all += cnt;
            tile_enum(T, d.r, w0, L, st, [&](int s, bool rev) {
                const int p = T.spath[s];
                ops.apply(d, a, tile_code(T, s, rev), T.p0 + p, m.seq_len, ratio_of(T.pU[p], total, cnt), 1);
            });
        }
        if (m.na > 1) st.tuples += n_all;                     // pass 1 deferred this mate without matching it
        code[single] = p2_single_outcome(sc, n_all);
    }
    d.w.grp_code2[g] = (uint8_t)(code[0] | (code[1] << 4));
}

// ---- kernels ----------------------------------------------------------------------------------------------------
template <int PASS, int K>
__global__ void __launch_bounds__(kTT) k_tile_pass(Dev d, sfb_tiles t, const long long* __restrict__ U) {
    extern __shared__ __align__(16) unsigned char smem[];
    TileS T = tile_carve(smem, t);
    BlockCounters& sc = *T.sc;
    block_counters_zero(sc);
    TileStats st{0, 0, 0, 0, 0, 0};
    const int4* items = reinterpret_cast<const int4*>(t.items);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) T.misc[0] = (int)atomicAdd(d.w.cursor + (PASS == 1 ? 5 : 6), 1ull);
        __syncthreads();
        const int item = T.misc[0];
        if (item >= (int)t.n_items) break;
        const int4 it = items[item];
        tile_stage<PASS>(T, d, t, it.x, U);
        for (i64 g = (i64)it.y + threadIdx.x; g < (i64)it.z; g += kTT) {
            if (PASS == 1) tile_group1<K>(T, d, sc, g, st);
            else tile_group2<K>(T, d, sc, g, st);
        }
        __syncthreads();
        tile_flush<PASS>(T, d, sc);
    }
    const int c0 = warp_sum(st.cand), c1 = warp_sum(st.cmp), c2 = warp_sum(st.tuples), c3 = warp_sum(st.rec),
              c4 = warp_sum(st.app), c5 = warp_sum(st.deferred);
    if (lane_id() == 0) {
        if (PASS == 1) {               // the matches of pass 2 are the same ones again: counted once
            SFB_COUNT(sc, SFB_C_ST_CAND, c0);
            SFB_COUNT(sc, SFB_C_ST_COMPARES, c1);
            SFB_COUNT(sc, SFB_C_ST_TUPLES, c2);
            SFB_COUNT(sc, SFB_C_ST_P1_RECORDS, c3);
            if (c5) atomicAdd(d.w.cursor + 7, (u64)c5);
        } else {
            SFB_COUNT(sc, SFB_C_ST_P2_RECORDS, c3);
            SFB_COUNT(sc, SFB_C_ST_P2_APPLIED, c4);
            SFB_COUNT(sc, SFB_C_ST_TUPLES, c2);
        }
    }
    __syncthreads();
    block_counters_flush(sc, d.acc.counters);
}

static int tile_check(const sfb_tiles* t) {
    SFB_REQUIRE(t, "tiles is null");
    SFB_REQUIRE(t->n_tiles >= 0 && t->n_items >= 0 && t->n_items < 0x7fffffffLL, "bad tile / item count");
    SFB_REQUIRE(t->tuple_cap == 8 || t->tuple_cap == 16 || t->tuple_cap == 32, "tuple_cap must be 8, 16 or 32");
    SFB_REQUIRE(t->max_slots >= 0 && t->max_slots <= 32767 && t->max_paths >= 0 && t->max_paths <= 255 &&
                    t->max_nodes >= 0 && t->max_nodes <= 65520, "tile exceeds the limits (32767 slots, 255 paths, 65520 nodes)");
    SFB_REQUIRE(t->n_items == 0 || (t->tile_desc && t->items && t->n_tiles > 0), "null tile array");
    return SFB_OK;
}

i64 tile_smem_bytes(const sfb_tiles* t) {
    if (tile_check(t) != SFB_OK) return SFB_E_ARG;
    return (i64)tile_layout((int)t->max_slots, (int)t->max_nodes, (int)t->max_paths, (int)t->tuple_cap).total;
}

template <int PASS, int K>
static int launch_tile_k(const Dev& d, const sfb_tiles* t, const long long* U, cudaStream_t st) {
    const size_t smem = tile_layout((int)t->max_slots, (int)t->max_nodes, (int)t->max_paths, K).total;
    auto kern = k_tile_pass<PASS, K>;
    if (smem > 227 * 1024) return fail(SFB_E_ARG, "tile kernels: %zu bytes of shared memory needed, 232448 available", smem);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return check_launch("k_tile_pass (shared memory size)");
    int per_sm = 0, dev = 0, sms = kSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTT, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const i64 grid = t->n_items < (i64)per_sm * sms ? t->n_items : (i64)per_sm * sms;
    kern<<<(unsigned)grid, kTT, smem, st>>>(d, *t, U);
    return check_launch(PASS == 1 ? "k_tile_pass1" : "k_tile_pass2");
}

int launch_tile_pass(int pass, const Dev& d, const sfb_tiles* t, const long long* U, cudaStream_t st) {
    SFB_TRY(tile_check(t));
    if (t->n_items == 0) return SFB_OK;
    if (pass == 1) {
        if (t->tuple_cap == 8) return launch_tile_k<1, 8>(d, t, U, st);
        if (t->tuple_cap == 16) return launch_tile_k<1, 16>(d, t, U, st);
        return launch_tile_k<1, 32>(d, t, U, st);
    }
    if (t->tuple_cap == 8) return launch_tile_k<2, 8>(d, t, U, st);
    if (t->tuple_cap == 16) return launch_tile_k<2, 16>(d, t, U, st);
    return launch_tile_k<2, 32>(d, t, U, st);
}

}  // namespace sfb
