// Tile kernels: the abundance pass for read groups that live inside one tile of the graph (sfb_tiles in sfb200.h).
//
// A thread block takes a work item (tile, range of read groups in tile order), stages the tile in shared memory --
// path nodes, node -> occurrence index, slot -> path, step index, per-path constants -- together with zeroed
// accumulator bins, and works through the item in rounds.  A round is a sequence of phases, each over a dense list so
// that the lanes of a warp stay busy (one thread per read group for everything ran at 4-8 active lanes of 32):
//   k_tile_pass1   M  thread = alignment of a window of 256 groups: get_matching_path (json2csv.py:341-363) on both
//                     orientations against the shared-memory index
//                  C  thread = read group: the pass-1 decision tree (:828-1053); unique assignments go to a list
//                  S  thread = unique assignment: fill_abund_dist (:321-339) -- per-path counters, histogram, node loop
//   k_tile_pass2   D  thread = read group, window after window until ~256 deferred groups are listed
//                  M  thread = listed alignment: the matches again (as the reference recomputes them, :1103-1114)
//                  C  thread = deferred group: candidate selection and ratios (:1123-1274); work items go to a list
//                  S  thread = work item: per-path sums (:1155-1156) and the node loop (:1157-1158)
// Matches: in a "simple" tile (at most 64 paths, no node twice in a path) an alignment's matches are two 64-bit path
// sets, the AND over the walk's steps of the step index, and the set logic of the decision tree is AND / OR / popcount;
// other tiles keep per-alignment tuple lists and run the generic logic of group_logic.cuh.
// Sums: shared-memory bins in 96-bit fixed point (three 32-bit limbs per accumulator, native ATOMS.ADD with exact
// carries: integer sums do not depend on the order of the additions), flushed once per work item with 64-bit integer
// REDs; histogram updates are aggregated per block in a small hash table keyed by (strain, five bins).
#include "common.cuh"
#include "bins.cuh"
#include "group_logic.cuh"

namespace sfb {

constexpr int kTT = 256;              // threads per block = read groups per window
constexpr int kRows1 = 3 * kTT;       // pass 1: alignments per window (a window shrinks until its alignments fit)
constexpr int kWin2 = 2 * kTT;        // pass 2: alignments per window ...
constexpr int kRows2 = 4 * kTT;       // ... listed alignments per round
constexpr int kGrp2 = 2 * kTT;        // ... listed groups per round
constexpr int kHash = 256;            // entries of the histogram aggregation table
constexpr u64 kHashEmpty = ~0ull;
constexpr unsigned kNoNode = 0xFFFEu; // local id of a walk node outside the tile (never equal to a path node)
constexpr unsigned kSentinel = 0xFFFFu;
constexpr int kOver = 255;            // tuple count of an alignment with more than KA matches
constexpr uint32_t kHasSlot = 0x80000000u;   // list entry carries its slot (tuple tiles); else phase S looks it up

__host__ __device__ inline int tile_rows(int pass) { return pass == 1 ? kRows1 : kRows2; }
__host__ __device__ inline int tile_list_cap(int pass) { return pass == 1 ? kRows1 + kTT : 6 * kTT; }   // entries of phase S's list
constexpr int kStageRec = 16 * kTT;   // pass 1: walk records of a window staged in shared memory (local node ids)
constexpr int kRec2 = 2 * kTT;        // pass 2: share records per round (total and count of a candidate set)
__host__ __device__ inline int tile_reps(int mp) { return mp <= 32 ? 8 : mp <= 64 ? 4 : 1; }       // copies of the per-path sums

// ---- shared-memory layout, sized by the largest tile (same arithmetic on host and device) -----------------------
struct TileLayout {
    unsigned hkey, pseq, pU, rec, emask, nmask, store, bins, plimb, la, lx, lseq, alist, glist, pclu, pstr0, pnstr, hcnt, sc, misc,
        lt, lm, grow, woff, wnode, goff, poff, node, occ, occ_off, eoff, esucc, spath, cnt, total;
};
__host__ __device__ inline TileLayout tile_layout(int pass, int ms, int mn, int mp, int me, int KA) {
    TileLayout L;
    unsigned o = 0;
    auto take = [&](unsigned bytes) { const unsigned at = o; o += (bytes + 15u) & ~15u; return at; };
    const unsigned rows = tile_rows(pass), cap = tile_list_cap(pass);
    const unsigned store_tup = 2u * rows * KA, store_mask = 16u * rows;
    L.hkey = take(8u * kHash);
    L.pseq = take(8u * mp);
    L.pU = take(pass == 2 ? 8u * mp : 0);
    L.rec = take(pass == 2 ? 16u * kRec2 : 0);  // pass 2: share records (long long total | ratio bits, int n, int len(sequence))
    L.emask = take(8u * me);                    // step index of a simple tile: path set of every step ...
    L.nmask = take(8u * mn);                    // ... and of every node
    L.store = take(store_tup > store_mask ? store_tup : store_mask);   // per alignment: match tuples (rev << 15 | slot),
                                                                       // or the forward / reverse path sets (simple tile)
    L.bins = take(12u * ms);
    L.plimb = take(28u * mp * tile_reps(mp));
    L.la = take(4u * cap);                      // list of phase S: alignment (offset inside the work item) ...
    L.lx = take(4u * cap);                      // ... path << 16 | reverse (the slot is looked up in phase S), or
                                                //     kHasSlot | path << 16 | slot ...
    L.lseq = take(pass == 1 ? 4u * cap : 0);    // ... len(sequence) of the mate (pass 1)
    L.alist = take(pass == 2 ? 4u * kRows2 : 0);   // pass 2: listed alignments (offset inside the work item)
    L.glist = take(pass == 2 ? 4u * kGrp2 : 0);    // pass 2: listed groups (offset inside the work item)
    L.pclu = take(4u * mp);
    L.pstr0 = take(4u * mp);
    L.pnstr = take(4u * mp);
    L.hcnt = take(4u * kHash);
    L.sc = take(sizeof(BlockCounters));
    L.misc = take(32);
    L.lt = take(pass == 2 ? 2u * cap : 0);      // list of phase S: times (pass 2) ...
    L.lm = take(pass == 2 ? 2u * cap : 0);      // ... and share record
    L.grow = take(pass == 2 ? 2u * kGrp2 : 0);  // pass 2: first row of a listed group
    L.woff = take(pass == 1 ? 4u * (kRows1 + 1) : 0);   // pass 1: walk offsets of the window's alignments (from its first record)
    L.wnode = take(pass == 1 ? 2u * kStageRec : 0);     // pass 1: local node ids of the window's first kStageRec records
    L.goff = take(4u * (kTT + 1));                      // alignment offsets of the next kTT groups (from the first one)
    L.poff = take(2u * (mp + 1));
    L.node = take(2u * (ms + 2));
    L.occ = take(2u * ms);
    L.occ_off = take(2u * (mn + 1));
    L.eoff = take(2u * (mn + 1));
    L.esucc = take(2u * me);
    L.spath = take(ms);
    L.cnt = take(rows);                         // tuples per alignment, kOver = more than KA
    L.total = o;
    return L;
}

// The block's shared memory.  Everything in it is addressed as (this array + an offset of the layout): the layout is a
// __grid_constant__ kernel parameter, so the offsets are constant-bank operands and every access is an LDS / STS / ATOMS
// on a known shared address (pointers kept in a struct ended up in local memory and turned every access into a generic
// load behind a local one).
extern __shared__ __align__(16) unsigned char sfb_smem[];

struct TileS {
    TileLayout L;        // by value: its fields are kernel constants (constant-bank operands), not loads through a pointer
    __device__ __forceinline__ u64* hkey() const { return reinterpret_cast<u64*>(sfb_smem + L.hkey); }
    __device__ __forceinline__ long long* pseq() const { return reinterpret_cast<long long*>(sfb_smem + L.pseq); }
    __device__ __forceinline__ long long* pU() const { return reinterpret_cast<long long*>(sfb_smem + L.pU); }
    __device__ __forceinline__ longlong2* rec() const { return reinterpret_cast<longlong2*>(sfb_smem + L.rec); }
    __device__ __forceinline__ u64* emask() const { return reinterpret_cast<u64*>(sfb_smem + L.emask); }
    __device__ __forceinline__ u64* nmask() const { return reinterpret_cast<u64*>(sfb_smem + L.nmask); }
    __device__ __forceinline__ u64* amask() const { return reinterpret_cast<u64*>(sfb_smem + L.store); }
    __device__ __forceinline__ uint16_t* tup() const { return reinterpret_cast<uint16_t*>(sfb_smem + L.store); }
    __device__ __forceinline__ uint32_t* bins() const { return reinterpret_cast<uint32_t*>(sfb_smem + L.bins); }
    __device__ __forceinline__ uint32_t* plimb() const { return reinterpret_cast<uint32_t*>(sfb_smem + L.plimb); }
    __device__ __forceinline__ uint32_t* la() const { return reinterpret_cast<uint32_t*>(sfb_smem + L.la); }
    __device__ __forceinline__ uint32_t* lx() const { return reinterpret_cast<uint32_t*>(sfb_smem + L.lx); }
    __device__ __forceinline__ int* lseq() const { return reinterpret_cast<int*>(sfb_smem + L.lseq); }
    __device__ __forceinline__ uint32_t* alist() const { return reinterpret_cast<uint32_t*>(sfb_smem + L.alist); }
    __device__ __forceinline__ uint32_t* glist() const { return reinterpret_cast<uint32_t*>(sfb_smem + L.glist); }
    __device__ __forceinline__ int* pclu() const { return reinterpret_cast<int*>(sfb_smem + L.pclu); }
    __device__ __forceinline__ int* pstr0() const { return reinterpret_cast<int*>(sfb_smem + L.pstr0); }
    __device__ __forceinline__ int* pnstr() const { return reinterpret_cast<int*>(sfb_smem + L.pnstr); }
    __device__ __forceinline__ uint32_t* hcnt() const { return reinterpret_cast<uint32_t*>(sfb_smem + L.hcnt); }
    __device__ __forceinline__ int* misc() const { return reinterpret_cast<int*>(sfb_smem + L.misc); }
    __device__ __forceinline__ uint16_t* lt() const { return reinterpret_cast<uint16_t*>(sfb_smem + L.lt); }
    __device__ __forceinline__ uint16_t* lm() const { return reinterpret_cast<uint16_t*>(sfb_smem + L.lm); }
    __device__ __forceinline__ uint16_t* grow() const { return reinterpret_cast<uint16_t*>(sfb_smem + L.grow); }
    __device__ __forceinline__ uint32_t* woff() const { return reinterpret_cast<uint32_t*>(sfb_smem + L.woff); }
    __device__ __forceinline__ uint16_t* wnode() const { return reinterpret_cast<uint16_t*>(sfb_smem + L.wnode); }
    __device__ __forceinline__ uint32_t* goff() const { return reinterpret_cast<uint32_t*>(sfb_smem + L.goff); }
    __device__ __forceinline__ uint16_t* poff() const { return reinterpret_cast<uint16_t*>(sfb_smem + L.poff); }
    __device__ __forceinline__ uint16_t* node() const { return reinterpret_cast<uint16_t*>(sfb_smem + L.node); }
    __device__ __forceinline__ uint16_t* occ() const { return reinterpret_cast<uint16_t*>(sfb_smem + L.occ); }
    __device__ __forceinline__ uint16_t* occ_off() const { return reinterpret_cast<uint16_t*>(sfb_smem + L.occ_off); }
    __device__ __forceinline__ uint16_t* eoff() const { return reinterpret_cast<uint16_t*>(sfb_smem + L.eoff); }
    __device__ __forceinline__ uint16_t* esucc() const { return reinterpret_cast<uint16_t*>(sfb_smem + L.esucc); }
    __device__ __forceinline__ uint8_t* spath() const { return reinterpret_cast<uint8_t*>(sfb_smem + L.spath); }
    __device__ __forceinline__ uint8_t* cnt() const { return reinterpret_cast<uint8_t*>(sfb_smem + L.cnt); }
    __device__ __forceinline__ BlockCounters* sc() const { return reinterpret_cast<BlockCounters*>(sfb_smem + L.sc); }
    int p0, np, n0, nn, s0, ns, mp, reps;
    bool simple;
    i64 a_item;          // first alignment of the work item
    i64 a_row0;          // pass 1: alignment of row 0 of the window
    i64 w_row0;          // pass 1: first walk record of the window
};
__device__ __forceinline__ TileS tile_carve(const TileLayout& L, const sfb_tiles& t) {
    TileS T;
    T.L = L;
    T.p0 = T.np = T.n0 = T.nn = T.s0 = T.ns = 0;
    T.mp = (int)t.max_paths;
    T.reps = tile_reps(T.mp);
    T.simple = false;
    T.a_item = T.a_row0 = T.w_row0 = 0;
    return T;
}

struct TileStats {
    int cand, cmp, tuples, rec, app, deferred;
};

// The node loop: cov[i] * weight into the bins of slot + i.  Records that cover their node (aligned == node length)
// add the weight itself, split once; with weight 1 that is the integer 1.
__device__ __forceinline__ void tile_add_records(const TileS& T, const sfb_reads& r, i64 w0, int L, int slot, bool unit,
                                                 double weight) {
    uint32_t* b = T.bins() + 3 * slot;
    Fx whole;
    whole.x0 = whole.x1 = 0u;
    whole.x2 = 1u;
    if (!unit) whole = fx_split(weight);
    for (int i = 0; i < L; ++i, b += 3) {
        const int v = r.walk_alen[w0 + i];
        if (v >= 0 && (v & 0xffff) == (v >> 16)) {
            if (unit) atomicAdd(b + 2, 1u); else fx_smem_add(b, whole);
            continue;
        }
        const double cov = v < 0 ? r.exc_cov[-(i64)v - 1] : (double)(v & 0xffff) / (double)(v >> 16);
        fx_smem_add(b, fx_split(unit ? cov : cov * weight));
    }
}

// ---- staging / flushing of a tile -------------------------------------------------------------------------------
template <int PASS>
__device__ __forceinline__ void tile_stage(TileS& T, const Dev& d, const sfb_tiles& t, int tile, const long long* U) {
    const int4 desc = reinterpret_cast<const int4*>(t.tile_desc)[2 * tile];
    T.simple = (reinterpret_cast<const int4*>(t.tile_desc)[2 * tile + 1].x & 1) != 0;
    T.p0 = desc.x;
    T.np = desc.y - desc.x;
    T.n0 = desc.z;
    T.nn = desc.w - desc.z;
    T.s0 = d.g.path_off[T.p0];
    T.ns = d.g.path_off[desc.y] - T.s0;
    const int tid = threadIdx.x;
    for (int i = tid; i < T.ns; i += kTT) {
        const int v = d.g.slot_node[T.s0 + i];
        T.node()[i] = v < 0 ? (uint16_t)kSentinel : (uint16_t)(v - T.n0);
    }
    if (tid < 2) T.node()[T.ns + tid] = (uint16_t)kSentinel;
    const int obase = d.g.occ_off[T.n0];
    for (int i = tid; i <= T.nn; i += kTT) T.occ_off()[i] = (uint16_t)(d.g.occ_off[T.n0 + i] - obase);
    const int nocc = d.g.occ_off[T.n0 + T.nn] - obase;
    const int2* occ = reinterpret_cast<const int2*>(d.g.occ_pair);
    for (int i = tid; i < nocc; i += kTT) T.occ()[i] = (uint16_t)(occ[obase + i].x - T.s0);
    for (int p = tid; p <= T.np; p += kTT) T.poff()[p] = (uint16_t)(d.g.path_off[T.p0 + p] - T.s0);
    if (T.simple) {
        const int ebase = t.edge_off[T.n0];
        for (int i = tid; i <= T.nn; i += kTT) T.eoff()[i] = (uint16_t)(t.edge_off[T.n0 + i] - ebase);
        const int ne = t.edge_off[T.n0 + T.nn] - ebase;
        for (int i = tid; i < ne; i += kTT) {
            T.esucc()[i] = (uint16_t)(t.edge_succ[ebase + i] - T.n0);
            T.emask()[i] = t.edge_mask[ebase + i];
        }
        for (int i = tid; i < T.nn; i += kTT) T.nmask()[i] = t.node_mask[T.n0 + i];
    }
    for (int p = tid; p < T.np; p += kTT) {
        T.pseq()[p] = d.g.path_seq_len[T.p0 + p];
        T.pclu()[p] = d.g.path_cluster[T.p0 + p];
        T.pstr0()[p] = d.g.path_strain0[T.p0 + p];
        T.pnstr()[p] = d.g.path_nstrain[T.p0 + p];
        if (PASS == 2) T.pU()[p] = U[T.p0 + p];
    }
    for (int i = tid; i < 3 * T.ns; i += kTT) T.bins()[i] = 0u;
    for (int i = tid; i < 7 * T.mp * T.reps; i += kTT) T.plimb()[i] = 0u;
    for (int i = tid; i < kHash; i += kTT) {
        T.hkey()[i] = kHashEmpty;
        T.hcnt()[i] = 0u;
    }
    __syncthreads();
    // slot -> local path: binary search over the staged path offsets
    for (int i = tid; i < T.ns; i += kTT) {
        int lo = 0, hi = T.np;               // largest p with poff[p] <= i
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if ((int)T.poff()[mid] <= i) lo = mid; else hi = mid;
        }
        T.spath()[i] = (uint8_t)lo;
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
// histogram update of a read on path `p` (local), aggregated in the block's table (json2csv.py:333-339 / :1246-1252)
__device__ __forceinline__ void tile_hist(const TileS& T, const Dev& d, BlockCounters& sc, int p, i64 a) {
    if (T.pnstr()[p] != 1) return;
    const uint8_t* b = d.r.aln_bins + a * SFB_N_HIST;
    const u64 key = ((u64)(uint32_t)T.pstr0()[p] << 40) | (u64)b[0] | ((u64)b[1] << 8) | ((u64)b[2] << 16) |
                    ((u64)b[3] << 24) | ((u64)b[4] << 32);
    unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 56) & (kHash - 1);
    for (int probe = 0; probe < kHash; ++probe, h = (h + 1) & (kHash - 1)) {
        const u64 old = atomicCAS(&T.hkey()[h], kHashEmpty, key);
        if (old == kHashEmpty || old == key) {
            atomicAdd(&T.hcnt()[h], 1u);
            return;
        }
    }
    hist_out(d, sc, key, 1u);                            // table full: straight to the global histograms
}

template <int PASS>
__device__ __forceinline__ void tile_flush(const TileS& T, const Dev& d, BlockCounters& sc) {
    const int tid = threadIdx.x;
    double* slots = PASS == 1 ? d.acc.uniq_abund : d.acc.mult_abund;
    for (int i = tid; i < T.ns; i += kTT) fx_flush(slots + T.s0 + i, d.acc.det_stride, T.bins() + 3 * i);
    for (int q = tid; q < T.np * T.reps; q += kTT) {
        const int p = q % T.np, rep = q / T.np;
        const uint32_t* l = T.plimb() + 7 * (rep * T.mp + p);
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
        if (T.hkey()[i] != kHashEmpty) hist_out(d, sc, T.hkey()[i], T.hcnt()[i]);
}
__device__ __forceinline__ uint32_t* tile_plimb(const TileS& T, int p) {      // this warp's copy of path p's sums
    return T.plimb() + 7 * (((threadIdx.x >> 5) % T.reps) * T.mp + p);
}

// ---- matching against the resident tile -------------------------------------------------------------------------
__device__ __forceinline__ unsigned tile_local(const TileS& T, int node) {
    const unsigned x = (unsigned)(node - T.n0);
    return x < (unsigned)T.nn ? x : kNoNode;
}
__device__ __forceinline__ uint32_t tile_code(const TileS& T, int slot, bool rev) {
    return (uint32_t)(T.s0 + slot) | (rev ? SFB_MATCH_REV : 0u);
}
// get_matching_path (json2csv.py:341-363) for the walk and, if longer than one node, the reversed walk: f(slot, rev)
// for every match, in the reference's order (forward by ascending slot, then reverse).  Any tile.
template <class F>
__device__ __forceinline__ void tile_enum(const TileS& T, const sfb_reads& r, i64 w0, int L, TileStats& st, F&& f) {
    if (L <= 0) return;
    const unsigned first = tile_local(T, r.walk_node[w0]);
    if (L == 1) {
        if (first == kNoNode) return;
        for (int c = T.occ_off()[first]; c < (int)T.occ_off()[first + 1]; ++c) {
            st.cand += 1;
            f((int)T.occ()[c], false);
        }
        return;
    }
    const unsigned last = tile_local(T, r.walk_node[w0 + L - 1]);
    const unsigned second = tile_local(T, r.walk_node[w0 + 1]), penult = tile_local(T, r.walk_node[w0 + L - 2]);
    if (first != kNoNode)
        for (int c = T.occ_off()[first]; c < (int)T.occ_off()[first + 1]; ++c) {
            const int s = T.occ()[c];
            st.cand += 1;
            st.cmp += 1;
            if (T.node()[s + 1] != second) continue;
            bool ok = true;
            for (int i = 2; i < L && ok; ++i) {
                st.cmp += 1;
                ok = T.node()[s + i] == tile_local(T, r.walk_node[w0 + i]);     // a sentinel ends the compare
            }
            if (ok) f(s, false);
        }
    if (last != kNoNode)
        for (int c = T.occ_off()[last]; c < (int)T.occ_off()[last + 1]; ++c) {
            const int s = T.occ()[c];
            st.cand += 1;
            st.cmp += 1;
            if (T.node()[s + 1] != penult) continue;
            bool ok = true;
            for (int i = 2; i < L && ok; ++i) {
                st.cmp += 1;
                ok = T.node()[s + i] == tile_local(T, r.walk_node[w0 + L - 1 - i]);
            }
            if (ok) f(s, true);
        }
}

// phase M, any tile: the matches of alignment `a` into row `row` of the tuple store (rev << 15 | slot)
template <int KA>
__device__ __forceinline__ void tile_match_aln(const TileS& T, const sfb_reads& r, int row, i64 a, TileStats& st) {
    const i64 w0 = r.aln_walk_off[a];
    const int L = (int)(r.aln_walk_off[a + 1] - w0);
    uint16_t* out = T.tup() + row * KA;
    int n = 0;
    tile_enum(T, r, w0, L, st, [&](int s, bool rev) {
        if (n < KA) out[n] = (uint16_t)(s | ((int)rev << 15));
        ++n;
    });
    T.cnt()[row] = (uint8_t)(n > KA ? kOver : n);
    st.tuples += n;
}

// The match tuples of a read group ("colored_match", json2csv.py:851-866) as stored by phase M: tuple i of mate k is
// found by walking the mate's alignments (one or two almost always).
template <int KA>
struct TileView {
    const TileS* T;
    int row[2];      // row of the mate's first alignment
    int na[2];
    int n[2];
    int c0[2];       // tuples of the mate's first alignment
    // false when an alignment overflowed its row
    __device__ __forceinline__ bool load(const TileS& T_, const Mate* mt, int row0) {
        T = &T_;
        bool fits = true;
        row[0] = row0;
        row[1] = row0 + mt[0].na;
        for (int k = 0; k < 2; ++k) {
            na[k] = mt[k].na;
            int tot = 0;
            for (int j = 0; j < na[k]; ++j) {
                const int c = T_.cnt()[row[k] + j];
                fits &= c != kOver;
                tot += c;
            }
            n[k] = tot;
            c0[k] = na[k] ? T_.cnt()[row[k]] : 0;
        }
        return fits;
    }
    __device__ __forceinline__ uint32_t word(int k, int i, int& j) const {
        j = 0;
        if (i >= c0[k]) {
            i -= c0[k];
            j = 1;
            for (int c; i >= (c = T->cnt()[row[k] + j]); ++j) i -= c;
        }
        return T->tup()[(row[k] + j) * KA + i];
    }
    __device__ __forceinline__ int nt(int k) const { return n[k]; }
    __device__ __forceinline__ int pid(int k, int i) const { int j; return T->p0 + T->spath()[word(k, i, j) & 0x7fffu]; }
    __device__ __forceinline__ uint32_t code(int k, int i) const {
        int j;
        const uint32_t w = word(k, i, j);
        return tile_code(*T, (int)(w & 0x7fffu), (w >> 15) != 0);
    }
    __device__ __forceinline__ int aloc(int k, int i) const { int j; word(k, i, j); return j; }
};

// ---- effects of the decisions (the Ops of group_logic.cuh) ---------------------------------------------------------
// Phase C only decides; what a decision adds to the accumulators is listed for phase S, where every lane has one entry.
// An entry names the alignment, the path and either the slot (tuple tiles) or the orientation (simple tiles: the slot
// is looked up in phase S); in pass 2 also the share record (total and count of the candidate set) and the
// multiplicity of the path (Q4).
struct TileOps {
    TileS* T;
    TileStats* st;
    int cap;
    __device__ __forceinline__ void hist(const Dev& d, BlockCounters& sc, int pid, i64 a) { tile_hist(*T, d, sc, pid - T->p0, a); }
    // fill_abund_dist (json2csv.py:321-339): listed with (tuple tiles) or without (simple tiles) the slot
    __device__ __forceinline__ void list_unique(i64 a, uint32_t x, int seq_len) {
        const int q = atomicAdd(&T->misc()[1], 1);             // an alignment is assigned at most once: q < kRows1
        T->la()[q] = (uint32_t)(a - T->a_item);
        T->lx()[q] = x;
        T->lseq()[q] = seq_len;
    }
    __device__ __forceinline__ void assign_unique(const Dev& d, BlockCounters& sc, i64 a, uint32_t code, int pid, int seq_len) {
        d.w.aln_assign[a] = (int32_t)code;
        list_unique(a, kHasSlot | ((uint32_t)(pid - T->p0) << 16) | (uint32_t)((int)(code & SFB_SLOT_MASK) - T->s0), seq_len);
    }
    // cluster of the first path through the walk's first node (json2csv.py:875,901,1036); -1 = None, -2 = no path
    __device__ __forceinline__ int first_cluster(const Dev& d, i64 a) {
        const unsigned n = tile_local(*T, d.r.walk_node[d.r.aln_walk_off[a]]);
        if (n == kNoNode || T->occ_off()[n] == T->occ_off()[n + 1]) return -2;
        return T->pclu()[T->spath()[T->occ()[T->occ_off()[n]]]];
    }
    __device__ __forceinline__ void book(const Dev& d, BlockCounters& sc, const Mate& mt, int cluster) {
        if (cluster < 0) { SFB_COUNT(sc, SFB_C_NO_CLUSTER, 1); return; }
        d.w.aln_assign[mt.a0] = -(cluster + 2);
    }
    __device__ __forceinline__ long long uniq(int pid) const { return T->pU()[pid - T->p0]; }
    // a share record: the shares of a candidate set are uniq[p] / total, or 1 / n when total is 0 (json2csv.py:1151-1154);
    // n < 0: `total` holds the bits of an explicit share.  -1 when the table is full.
    __device__ __forceinline__ int record(long long total, int n, int seq_len) {
        const int k = atomicAdd(&T->misc()[4], 1);
        if (k >= kRec2) return -1;
        T->rec()[k] = make_longlong2(total, (long long)(uint32_t)n | ((long long)seq_len << 32));
        return k;
    }
    // one work item of pass 2 (json2csv.py:1155-1158 and twins); false when a list is full
    __device__ __forceinline__ bool list_share(i64 a, uint32_t x, int times, int rec) {
        st->app += 1;
        if (rec < 0 || times > 0xffff) return false;
        const int q = atomicAdd(&T->misc()[1], 1);
        if (q >= cap) return false;
        T->la()[q] = (uint32_t)(a - T->a_item);
        T->lx()[q] = x;
        T->lt()[q] = (uint16_t)times;
        T->lm()[q] = (uint16_t)rec;
        return true;
    }
    // generic logic (tuple tiles): the share comes computed
    __device__ __forceinline__ void apply(const Dev& d, i64 a, uint32_t code, int pid, int seq_len, double ratio, int times) {
        const int slot = (int)(code & SFB_SLOT_MASK) - T->s0, p = pid - T->p0;
        if (list_share(a, kHasSlot | ((uint32_t)p << 16) | (uint32_t)slot, times, record(__double_as_longlong(ratio), -1, seq_len)))
            return;
        st->app -= 1;
        apply_now(d, a, slot, p, seq_len, ratio, times);
    }
    // a list is full (many candidate paths per read): here and now
    __device__ __forceinline__ void apply_now(const Dev& d, i64 a, int slot, int p, int seq_len, double ratio, int times) {
        st->app += 1;
        const double weight = ratio * (double)times;
        uint32_t* l = tile_plimb(*T, p);
        l[0] = 1u;                                           // "touched" flag
        fx_smem_add(l + 1, fx_split(weight));
        fx_smem_add(l + 4, fx_split((ratio * (double)seq_len / (double)T->pseq()[p]) * (double)times));
        const i64 w0 = d.r.aln_walk_off[a];
        const int L = (int)(d.r.aln_walk_off[a + 1] - w0);
        tile_add_records(*T, d.r, w0, L, slot, false, weight);
        st->rec += L;
    }
};
// slot of node `n` (local) on path `p` (local) of a simple tile
__device__ __forceinline__ int tile_slot_of(const TileS& T, unsigned n, int p) {
    for (int c = T.occ_off()[n]; c < (int)T.occ_off()[n + 1]; ++c)
        if (T.spath()[T.occ()[c]] == p) return T.occ()[c];
    return 0;
}
// slot of a list entry: carried, or (simple tile) where path p has the walk's first node -- its last node for a match of
// the reversed walk (json2csv.py:862)
__device__ __forceinline__ int tile_entry_slot(const TileS& T, const sfb_reads& r, uint32_t x, i64 w0, int L) {
    if (x & kHasSlot) return (int)(x & 0x7fffu);
    const i64 w = (x & 1u) ? w0 + L - 1 : w0;
    return tile_slot_of(T, tile_local(T, r.walk_node[w]), (int)((x >> 16) & 0xffu));
}
// phase S of pass 1: one unique assignment (fill_abund_dist, json2csv.py:321-339)
__device__ __forceinline__ void tile_scatter1(TileS& T, const Dev& d, BlockCounters& sc, int q, TileStats& st) {
    const i64 a = T.a_item + T.la()[q];
    const uint32_t x = T.lx()[q];
    const int p = (int)((x >> 16) & 0xffu);
    const i64 w0 = d.r.aln_walk_off[a];
    const int L = (int)(d.r.aln_walk_off[a + 1] - w0);
    const int slot = tile_entry_slot(T, d.r, x, w0, L);
    if (!(x & kHasSlot)) d.w.aln_assign[a] = (int32_t)tile_code(T, slot, (x & 1u) != 0);
    uint32_t* l = tile_plimb(T, p);
    atomicAdd(l, 1u);
    fx_smem_add(l + 1, fx_split((double)T.lseq()[q] / (double)T.pseq()[p]));
    tile_hist(T, d, sc, p, a);
    tile_add_records(T, d.r, w0, L, slot, true, 1.0);
    st.rec += L;
}
// phase S of pass 2: one work item (json2csv.py:1155-1158)
__device__ __forceinline__ void tile_scatter2(TileS& T, const Dev& d, int q, TileStats& st) {
    if (T.lt()[q] == 0) return;                                // reserved, then applied at once (a table was full)
    const i64 a = T.a_item + T.la()[q];
    const uint32_t x = T.lx()[q];
    const int p = (int)((x >> 16) & 0xffu);
    const longlong2 rc = T.rec()[T.lm()[q]];
    const int n = (int)(uint32_t)rc.y, seq_len = (int)(rc.y >> 32);
    const double ratio = n < 0 ? __longlong_as_double(rc.x) : ratio_of(T.pU()[p], rc.x, n);
    const double times = (double)T.lt()[q];
    const double weight = ratio * times;
    const i64 w0 = d.r.aln_walk_off[a];
    const int L = (int)(d.r.aln_walk_off[a + 1] - w0);
    const int slot = tile_entry_slot(T, d.r, x, w0, L);
    uint32_t* l = tile_plimb(T, p);
    l[0] = 1u;                                               // "touched" flag
    fx_smem_add(l + 1, fx_split(weight));
    fx_smem_add(l + 4, fx_split((ratio * (double)seq_len / (double)T.pseq()[p]) * times));
    tile_add_records(T, d.r, w0, L, slot, false, weight);
    st.rec += L;
}

// =================================================================================================================
// Simple tiles (at most 64 paths, no node twice in a path): path sets as 64-bit masks
// =================================================================================================================
// A walk matches path p (at the position of its first node) iff p takes every step of the walk; the reversed walk
// matches iff p takes every step backwards.  So the matches of an alignment are two path sets, the AND over the
// walk's steps of the step index, and the set logic of the decision tree becomes AND / OR / popcount on them.  The
// tuple order of the reference (json2csv.py:851-866: alignment by alignment, forward matches by ascending path, then
// reverse) only matters for picking "the first" tuple; sums are exact integers here, so their order is free.
__device__ __forceinline__ u64 tile_step(const TileS& T, unsigned n, unsigned succ) {
    const int e0 = T.eoff()[n], e1 = T.eoff()[n + 1];
    // almost every node has one or two successors: two predicated probes, no loop
    u64 m = 0;
    if (e0 < e1 && T.esucc()[e0] == succ) m = T.emask()[e0];
    if (e0 + 1 < e1 && T.esucc()[e0 + 1] == succ) m = T.emask()[e0 + 1];
    for (int e = e0 + 2; e < e1; ++e)
        if (T.esucc()[e] == succ) m = T.emask()[e];
    return m;
}
// phase M: forward / reverse path sets of alignment `a` into row `row`.  STAGED (pass 1): the walk offsets and the first
// kStageRec records of the window are in shared memory.
template <bool STAGED>
__device__ __forceinline__ void mask_match_aln(const TileS& T, const sfb_reads& r, int row, i64 a, TileStats& st) {
    i64 w0;
    int L;
    if (STAGED) {
        w0 = T.w_row0 + T.woff()[row];
        L = (int)(T.woff()[row + 1] - T.woff()[row]);
    } else {
        w0 = r.aln_walk_off[a];
        L = (int)(r.aln_walk_off[a + 1] - w0);
    }
    const unsigned rel = STAGED ? T.woff()[row] : 0u;
    auto node_at = [&](int k) -> unsigned {
        if (STAGED && rel + (unsigned)k < (unsigned)kStageRec) return T.wnode()[rel + k];
        return tile_local(T, r.walk_node[w0 + k]);
    };
    u64 f = 0, rv = 0;
    if (L > 0) {
        unsigned prev = node_at(0);
        if (L == 1) {
            if (prev != kNoNode) f = T.nmask()[prev];
        } else if (prev != kNoNode) {
            f = rv = ~0ull;
            for (int k = 1; k < L; ++k) {
                const unsigned n = node_at(k);
                if (n == kNoNode) { f = rv = 0; break; }
                if (f) f &= tile_step(T, prev, n);
                if (rv) rv &= tile_step(T, n, prev);
                st.cmp += 1;
                if (!(f | rv)) break;
                prev = n;
            }
        }
    }
    T.amask()[2 * row] = f;
    T.amask()[2 * row + 1] = rv;
    const int n = __popcll(f) + __popcll(rv);
    st.cand += n;
    st.tuples += n;
}
// match code of the tuple (alignment a, path p, orientation), for the rare paths that need it at once
__device__ __forceinline__ uint32_t mask_code(const TileS& T, const sfb_reads& r, i64 a, int p, bool rev) {
    const i64 w = rev ? r.aln_walk_off[a + 1] - 1 : r.aln_walk_off[a];
    return tile_code(T, tile_slot_of(T, tile_local(T, r.walk_node[w]), p), rev);
}

struct MaskMate {
    int row, na, nt;     // row of the first alignment, alignments, tuples
    u64 S, dup;          // paths matched, paths matched more than once
    __device__ __forceinline__ void load(const TileS& T, int row_, int na_) {
        row = row_;
        na = na_;
        S = dup = 0;
        nt = 0;
        for (int j = 0; j < 2 * na; ++j) {
            const u64 x = T.amask()[2 * row + j];
            dup |= S & x;
            S |= x;
            nt += __popcll(x);
        }
    }
    __device__ __forceinline__ int count(const TileS& T, int p) const {      // list_path_id.count(pid)
        int c = 0;
        for (int j = 0; j < 2 * na; ++j) c += (int)((T.amask()[2 * row + j] >> p) & 1ull);
        return c;
    }
    // the first tuple carrying path p: alignment j inside the mate, orientation
    __device__ __forceinline__ void find(const TileS& T, int p, int& j, bool& rev) const {
        for (int q = 0; q < 2 * na; ++q)
            if ((T.amask()[2 * row + q] >> p) & 1ull) { j = q >> 1; rev = q & 1; return; }
        j = 0;
        rev = false;
    }
};

// strains common to both mates' candidate paths, one word (json2csv.py:964-972)
__device__ __forceinline__ u64 mask_common(const sfb_graph& g, const TileS& T, const MaskMate* mm, int word) {
    u64 u[2] = {0, 0};
    for (int k = 0; k < 2; ++k)
        for (u64 rest = mm[k].S; rest; rest &= rest - 1)
            u[k] |= g.path_smask[(size_t)(T.p0 + __ffsll((long long)rest) - 1) * g.mask_words + word];
    return u[0] & u[1];
}
// in how many common strains path p occurs (Q4: a tuple counts once per common strain its path carries)
__device__ __forceinline__ int mask_times(const sfb_graph& g, const TileS& T, const MaskMate* mm, u64 common0, int p) {
    if (g.mask_words == 1) return __popcll(g.path_smask[T.p0 + p] & common0);
    int t = 0;
    for (int word = 0; word < (int)g.mask_words; ++word)
        t += __popcll(g.path_smask[(size_t)(T.p0 + p) * g.mask_words + word] & mask_common(g, T, mm, word));
    return t;
}
struct MaskCut {
    bool any;
    int nn[2];
    u64 common0;
};
__device__ __forceinline__ MaskCut mask_cut(const sfb_graph& g, const TileS& T, const MaskMate* mm) {
    MaskCut c;
    c.any = false;
    c.nn[0] = c.nn[1] = 0;
    c.common0 = 0;
    for (int word = 0; word < (int)g.mask_words; ++word) {
        const u64 common = mask_common(g, T, mm, word);
        if (word == 0) c.common0 = common;
        if (!common) continue;
        c.any = true;
        for (int k = 0; k < 2; ++k)
            for (u64 rest = mm[k].S; rest; rest &= rest - 1) {
                const int p = __ffsll((long long)rest) - 1;
                c.nn[k] += mm[k].count(T, p) * __popcll(g.path_smask[(size_t)(T.p0 + p) * g.mask_words + word] & common);
            }
    }
    return c;
}

// both mates have colored paths (json2csv.py:918-1000); same decisions as classify_pair (group_logic.cuh)
__device__ __forceinline__ void mask_classify_pair(TileS& T, const Dev& d, TileOps& ops, BlockCounters& sc, const Mate* mt,
                                                   const MaskMate* mm, int* code, bool& defer) {
    const u64 inter = mm[0].S & mm[1].S;
    const int n_inter = __popcll(inter);
    if (n_inter > 1) {
        SFB_COUNT(sc, SFB_C_SECOND_PASS, 2);
        defer = true;
        code[0] = code[1] = SFB_DEFER;
        return;
    }
    if (n_inter == 1) {
        const int p = __ffsll((long long)inter) - 1;
        for (int k = 0; k < 2; ++k) {
            if ((mm[k].dup >> p) & 1ull) {
                SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
                code[k] = SFB_MULTI_ALIGN;
                continue;
            }
            int j;
            bool rev;
            mm[k].find(T, p, j, rev);
            ops.list_unique(mt[k].a0 + j, ((uint32_t)p << 16) | (uint32_t)rev, mt[k].seq_len);
            SFB_COUNT(sc, SFB_C_ASSIGNED_U, 1);
            SFB_COUNT(sc, SFB_C_SAME_CP, 1);
            if (mm[k].nt > 1) SFB_COUNT(sc, SFB_C_AMBIG_CP_RESOLVED, 1);
            code[k] = SFB_ASSIGN_SAME;
        }
        return;
    }
    // no common path: intersect strains (json2csv.py:961-1000)
    const MaskCut cut = mask_cut(d.g, T, mm);
    if (!cut.any) {
        SFB_COUNT(sc, SFB_C_INCONSIST, 2);
        code[0] = code[1] = SFB_INCONSIST;
        return;
    }
    for (int k = 0; k < 2; ++k) {
        const int nn = cut.nn[k];
        if (nn < mm[k].nt) SFB_COUNT(sc, nn == 1 ? SFB_C_AMBIG_RESOLVED : SFB_C_AMBIG_REDUCED, 1);
        if (nn > 1) {
            SFB_COUNT(sc, SFB_C_SECOND_PASS, 1);
            defer = true;
            code[k] = SFB_DEFER;
            continue;
        }
        for (u64 rest = mm[k].S; rest; rest &= rest - 1) {       // exactly one tuple survives, once
            const int p = __ffsll((long long)rest) - 1;
            if (mask_times(d.g, T, mm, cut.common0, p) == 0) continue;
            if (mm[k].count(T, p) == 1) {
                int j;
                bool rev;
                mm[k].find(T, p, j, rev);
                ops.list_unique(mt[k].a0 + j, ((uint32_t)p << 16) | (uint32_t)rev, mt[k].seq_len);
                SFB_COUNT(sc, SFB_C_ASSIGNED_U, 1);
                SFB_COUNT(sc, SFB_C_DIFF_CP, 1);
                code[k] = SFB_ASSIGN_DIFF;
            } else {
                SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
                code[k] = SFB_MULTI_ALIGN;
            }
            break;
        }
    }
}

// every tuple of mate k whose path is in `sel` becomes a work item of share record `rec`, times(p) times
template <class W>
__device__ __forceinline__ void mask_list(TileS& T, const Dev& d, TileOps& ops, const Mate& m, int row, int n_masks, u64 sel,
                                          long long total, int n, W&& times_of_path) {
    int cnt = 0;
    for (int q = 0; q < n_masks; ++q) cnt += __popcll(T.amask()[2 * row + q] & sel);
    if (cnt == 0) return;
    const int rec = ops.record(total, n, m.seq_len);
    int at = atomicAdd(&T.misc()[1], cnt);             // one reservation per candidate set, not one per work item
    ops.st->app += cnt;
    for (int q = 0; q < n_masks; ++q)
        for (u64 rest = T.amask()[2 * row + q] & sel; rest; rest &= rest - 1, ++at) {
            const int p = __ffsll((long long)rest) - 1;
            const i64 a = m.a0 + (q >> 1);
            const int times = times_of_path(p);
            if (rec >= 0 && at < ops.cap && times <= 0xffff) {
                T.la()[at] = (uint32_t)(a - T.a_item);
                T.lx()[at] = ((uint32_t)p << 16) | (uint32_t)(q & 1);
                T.lt()[at] = (uint16_t)times;
                T.lm()[at] = (uint16_t)rec;
                continue;
            }
            ops.st->app -= 1;                // a list is full: the share here and now
            if (at < ops.cap) T.lt()[at] = 0;  // (a reserved entry that stays empty)
            ops.apply_now(d, a, (int)(mask_code(T, d.r, a, p, (q & 1) != 0) & SFB_SLOT_MASK) - T.s0, p, m.seq_len,
                          ratio_of(T.pU()[p], total, n), times);
        }
}

// pass 2 for an aligned pair whose mates both have colored paths (json2csv.py:1123-1217); same decisions as
// p2_decide_pair + p2_emit (group_logic.cuh)
__device__ __forceinline__ void mask_pass2_pair(TileS& T, const Dev& d, TileOps& ops, BlockCounters& sc, const Mate* mt,
                                                const MaskMate* mm, int* code) {
    const u64 inter = mm[0].S & mm[1].S;
    const int n_inter = __popcll(inter);
    if (n_inter >= 1) {
        long long total = 0;
        for (u64 rest = inter; rest; rest &= rest - 1) total += T.pU()[__ffsll((long long)rest) - 1];
        for (int k = 0; k < 2; ++k) {
            if (mm[k].dup & inter) {             // a common path with several tuples in this mate
                SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
                code[k] = SFB_P2_MULTI_ALIGN;
                continue;
            }
            SFB_COUNT(sc, SFB_C_ASSIGNED_M, 1);
            code[k] = SFB_P2_ASSIGNED;
            mask_list(T, d, ops, mt[k], mm[k].row, 2 * mm[k].na, inter, total, n_inter, [](int) { return 1; });
        }
        return;
    }
    const MaskCut cut = mask_cut(d.g, T, mm);
    if (!cut.any) {
        SFB_COUNT(sc, SFB_C_INCONSIST_M, (mm[0].nt == 1 || mm[1].nt == 1) ? 1 : 2);
        code[0] = code[1] = SFB_P2_INCONSIST;
        return;
    }
    for (int k = 0; k < 2; ++k) {
        const int nn = cut.nn[k];
        if (nn <= 1) continue;                   // handled in pass 1 (json2csv.py:1185)
        bool simple = true;
        long long tot = 0;
        u64 sel = 0;
        for (u64 rest = mm[k].S; rest && simple; rest &= rest - 1) {
            const int p = __ffsll((long long)rest) - 1;
            const int times = mask_times(d.g, T, mm, cut.common0, p);
            if (times == 0) continue;
            if (mm[k].count(T, p) != 1) { simple = false; break; }
            tot += T.pU()[p] * times;
            sel |= 1ull << p;
        }
        if (!simple) {
            SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
            code[k] = SFB_P2_MULTI_ALIGN;
            continue;
        }
        SFB_COUNT(sc, SFB_C_ASSIGNED_M, 1);
        code[k] = SFB_P2_ASSIGNED;
        mask_list(T, d, ops, mt[k], mm[k].row, 2 * mm[k].na, sel, tot, nn,
                  [&](int p) { return mask_times(d.g, T, mm, cut.common0, p); });
    }
}

// An aligned pair that does not fit the tuple buffers: its match lists in the format of sfb_match, the group on the
// list of the generic classify kernel (which also takes care of its node loops and of its deferral).
__device__ __forceinline__ void tile_export(const TileS& T, const Dev& d, BlockCounters& sc, i64 g, const Mate* mt,
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
                if (n++ == 0) only = make_uint2(tile_code(T, s, rev), (uint32_t)(T.p0 + T.spath()[s]));
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
                match_buf[pos + q++] = make_uint2(tile_code(T, s, rev), (uint32_t)(T.p0 + T.spath()[s]));
            });
            aln_match[a] = make_uint2(SFB_MATCH_MULTI | (uint32_t)pos, (uint32_t)n);
        }
    d.w.slow[atomicAdd(d.w.cursor + 3, 1ull)] = (int32_t)g;
}

// ---- phase C of pass 1, one read group (json2csv.py:828-1053); row0 = row of its first alignment ------------------
template <int KA>
__device__ __forceinline__ void tile_group1(TileS& T, const Dev& d, BlockCounters& sc, i64 g, int row0, TileStats& st) {
    const int nm = d.r.grp_nmates[g];
    Mate mt[2];
    mate_load(d.r, g, 0, mt[0]);
    mate_load(d.r, g, 1, mt[1]);
    int code[2] = {SFB_ABSENT, SFB_ABSENT};
    bool defer = false;
    SFB_COUNT(sc, SFB_C_NB_READS, nm);
    int single = nm == 1 ? 0 : -1;           // mate handled single-end style, -1 = none
    TileOps ops{&T, &st, 0};
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
            TileView<KA> v;
            MaskMate mm[2];
            int nt[2];
            if (T.simple) {
                mm[0].load(T, row0, mt[0].na);
                mm[1].load(T, row0 + mt[0].na, mt[1].na);
                nt[0] = mm[0].nt;
                nt[1] = mm[1].nt;
            } else {
                if (!v.load(T, mt, row0) || mt[0].na > 255 || mt[1].na > 255) {
                    tile_export(T, d, sc, g, mt, st);
                    return;
                }
                nt[0] = v.n[0];
                nt[1] = v.n[1];
            }
            if (nt[0] == 0 && nt[1] == 0) {
                classify_both_unassigned(d, ops, sc, mt, code);
            } else if (nt[0] == 0 || nt[1] == 0) {
                // one mate without colored path (json2csv.py:895-916)
                SFB_COUNT(sc, SFB_C_UNASSIGNED, 1);
                SFB_COUNT(sc, SFB_C_ONE_NOCP, 1);
                const int gone = nt[0] == 0 ? 0 : 1;
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
            } else if (T.simple) {
                mask_classify_pair(T, d, ops, sc, mt, mm, code, defer);
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
            const int row = row0 + (single ? mt[0].na : 0);
            const int n = T.simple ? __popcll(T.amask()[2 * row]) + __popcll(T.amask()[2 * row + 1])
                                   : T.cnt()[row];                 // kOver: more than KA, in any case more than one
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

// ---- phase C of pass 2, one deferred read group (json2csv.py:1085-1274) -----------------------------------------
template <int KA>
__device__ __forceinline__ void tile_group2(TileS& T, const Dev& d, BlockCounters& sc, i64 g, int row0, TileStats& st) {
    const int nm = d.r.grp_nmates[g];
    Mate mt[2];
    mate_load(d.r, g, 0, mt[0]);
    mate_load(d.r, g, 1, mt[1]);
    int code[2] = {SFB_P2_NONE, SFB_P2_NONE};
    int single = nm == 1 ? 0 : -1;
    TileOps ops{&T, &st, tile_list_cap(2)};
    if (nm == 2) {
        if (mt[0].na == 0 || mt[1].na == 0) {
            single = mt[0].na == 0 ? 1 : 0;                                  // json2csv.py:1088-1093
        } else {
            TileView<KA> v;
            MaskMate mm[2];
            int nt[2];
            if (T.simple) {
                mm[0].load(T, row0, mt[0].na);
                mm[1].load(T, row0 + mt[0].na, mt[1].na);
                nt[0] = mm[0].nt;
                nt[1] = mm[1].nt;
            } else {
                if (!v.load(T, mt, row0) || mt[0].na > 255 || mt[1].na > 255) return;  // handed over in pass 1: sfb_pass2 has it
                nt[0] = v.n[0];
                nt[1] = v.n[1];
            }
            if (nt[0] == 0 || nt[1] == 0) {
                single = nt[0] == 0 ? 1 : 0;                                 // :1117-1122
            } else if (T.simple) {
                mask_pass2_pair(T, d, ops, sc, mt, mm, code);
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
        // every alignment separately (json2csv.py:1230-1268)
        const Mate& m = mt[single];
        int n_all = 0;
        for (int j = 0; j < m.na; ++j) {
            const i64 a = m.a0 + j;
            const int row = row0 + (single ? mt[0].na : 0) + j;
            if (T.simple) {
                const u64 f = T.amask()[2 * row], rv = T.amask()[2 * row + 1];
                const int cnt = __popcll(f) + __popcll(rv);
                if (cnt == 0) continue;
                long long total = 0;
                for (int q = 0; q < 2; ++q)
                    for (u64 rest = q ? rv : f; rest; rest &= rest - 1) {
                        const int p = __ffsll((long long)rest) - 1;
                        total += T.pU()[p];
                        tile_hist(T, d, sc, p, a);                           // :1246-1252
                    }
                Mate one = m;                                                // this alignment alone
                one.a0 = a;
                mask_list(T, d, ops, one, row, 2, ~0ull, total, cnt, [](int) { return 1; });
                n_all += cnt;
                continue;
            }
            const int cnt = T.cnt()[row];
            if (cnt == kOver) {
                // more matches than a row holds: enumerated twice from the index instead of stored
                const i64 w0 = d.r.aln_walk_off[a];
                const int L = (int)(d.r.aln_walk_off[a + 1] - w0);
                long long total = 0;
                int c2 = 0;
                tile_enum(T, d.r, w0, L, st, [&](int s, bool) {
                    const int p = T.spath()[s];
                    total += T.pU()[p];
                    ++c2;
                    tile_hist(T, d, sc, p, a);
                });
                tile_enum(T, d.r, w0, L, st, [&](int s, bool rev) {
                    const int p = T.spath()[s];
                    ops.apply(d, a, tile_code(T, s, rev), T.p0 + p, m.seq_len, ratio_of(T.pU()[p], total, c2), 1);
                });
                n_all += c2;
                continue;
            }
            if (cnt == 0) continue;
            const uint16_t* tp = T.tup() + row * KA;
            long long total = 0;
            for (int i = 0; i < cnt; ++i) {
                const int p = T.spath()[tp[i] & 0x7fffu];
                total += T.pU()[p];
                tile_hist(T, d, sc, p, a);
            }
            for (int i = 0; i < cnt; ++i) {
                const int s = tp[i] & 0x7fffu, p = T.spath()[s];
                ops.apply(d, a, tile_code(T, s, (tp[i] >> 15) != 0), T.p0 + p, m.seq_len, ratio_of(T.pU()[p], total, cnt), 1);
            }
            n_all += cnt;
        }
        code[single] = p2_single_outcome(sc, n_all);
    }
    d.w.grp_code2[g] = (uint8_t)(code[0] | (code[1] << 4));
}

// ---- kernels ----------------------------------------------------------------------------------------------------
// End of the window of read groups that starts at g0: at most `groups` groups and at most `cap` alignments, trimmed to
// at most `want` alignments when want > 0 (whole iterations of phase M).  The offsets of the candidate groups are staged
// in shared memory with one coalesced load, the searches then run there.  Block-uniform; ends with a barrier.
__device__ __forceinline__ i64 tile_window(TileS& T, const sfb_reads& r, i64 g0, i64 g_end, int cap, int groups = kTT, int want = 0) {
    const int n = (int)min((i64)groups, g_end - g0);
    const i64 a0 = r.grp_aln_off[2 * g0];
    for (int i = threadIdx.x; i <= n; i += kTT) {
        const i64 v = r.grp_aln_off[2 * (g0 + i)] - a0;
        T.goff()[i] = (uint32_t)min(v, (i64)0x7fffffff);
    }
    __syncthreads();
    int lim = cap;
    if (want > 0 && want < cap && (int)T.goff()[1] <= want) lim = want;        // (a first group above `want` keeps `cap`)
    int lo = 1, hi = n;                      // largest k in [1, n] with goff[k] <= lim (k = 1 whatever the first group has)
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int)T.goff()[mid] <= lim) lo = mid; else hi = mid - 1;
    }
    return g0 + lo;
}
// the arrays of read groups [g0, g1) on their way to the L2 (they stream from HBM, every byte is used once)
__device__ __forceinline__ void tile_prefetch(const sfb_reads& r, i64 g0, i64 g1, i64 g_end) {
    g1 = min(g1, g_end);
    if (g0 >= g1) return;
    auto lines = [&](const void* base, size_t lo, size_t hi) {       // bytes [lo, hi) of an array, one line per thread
        const char* p = (const char*)base;
        for (size_t b = (lo & ~(size_t)127) + 128 * (size_t)threadIdx.x; b < hi; b += 128 * (size_t)kTT)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p + b));
    };
    lines(r.grp_aln_off, 16 * (size_t)g0, 16 * (size_t)g1 + 8);
    lines(r.grp_nmates, (size_t)g0, (size_t)g1);
    lines(r.mate_seq_len, 8 * (size_t)g0, 8 * (size_t)g1);
}
__device__ __forceinline__ void tile_prefetch_alns(const sfb_reads& r, i64 a0, i64 a1) {
    if (a0 >= a1) return;
    auto lines = [&](const void* base, size_t lo, size_t hi) {
        const char* p = (const char*)base;
        for (size_t b = (lo & ~(size_t)127) + 128 * (size_t)threadIdx.x; b < hi; b += 128 * (size_t)kTT)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p + b));
    };
    lines(r.aln_walk_off, 8 * (size_t)a0, 8 * (size_t)a1 + 8);
    lines(r.aln_bins, 5 * (size_t)a0, 5 * (size_t)a1);
}
__device__ __forceinline__ void tile_prefetch_recs(const sfb_reads& r, i64 w0, i64 w1) {
    if (w0 >= w1) return;
    auto lines = [&](const void* base, size_t lo, size_t hi) {
        const char* p = (const char*)base;
        for (size_t b = (lo & ~(size_t)127) + 128 * (size_t)threadIdx.x; b < hi; b += 128 * (size_t)kTT)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p + b));
    };
    lines(r.walk_node, 4 * (size_t)w0, 4 * (size_t)w1);
    lines(r.walk_alen, 4 * (size_t)w0, 4 * (size_t)w1);
}

#ifndef SFB_TILE_MINBLK
#define SFB_TILE_MINBLK 3
#endif
template <int PASS, int KA>
__global__ void __launch_bounds__(kTT, SFB_TILE_MINBLK) k_tile_pass(const __grid_constant__ Dev d, const __grid_constant__ sfb_tiles t,
                                                                        const __grid_constant__ TileLayout lay,
                                                                        const long long* __restrict__ U) {
    // every array of the call lives in global memory: say so, or each access is a generic load
#define SFB_GLOBAL(p) __builtin_assume(__isGlobal(p))
    SFB_GLOBAL(d.r.grp_nmates); SFB_GLOBAL(d.r.grp_aln_off); SFB_GLOBAL(d.r.mate_seq_len); SFB_GLOBAL(d.r.aln_walk_off);
    SFB_GLOBAL(d.r.aln_bins); SFB_GLOBAL(d.r.walk_node); SFB_GLOBAL(d.r.walk_alen); SFB_GLOBAL(d.r.exc_cov);
    SFB_GLOBAL(d.w.aln_match); SFB_GLOBAL(d.w.match_buf); SFB_GLOBAL(d.w.cursor); SFB_GLOBAL(d.w.grp_code);
    SFB_GLOBAL(d.w.grp_code2); SFB_GLOBAL(d.w.aln_assign); SFB_GLOBAL(d.w.slow);
    SFB_GLOBAL(d.g.slot_node); SFB_GLOBAL(d.g.path_off); SFB_GLOBAL(d.g.occ_off); SFB_GLOBAL(d.g.occ_pair);
    SFB_GLOBAL(d.g.path_seq_len); SFB_GLOBAL(d.g.path_nstrain); SFB_GLOBAL(d.g.path_strain0); SFB_GLOBAL(d.g.path_smask);
    SFB_GLOBAL(d.g.path_cluster);
    SFB_GLOBAL(d.acc.uniq_reads); SFB_GLOBAL(d.acc.uniq_norm); SFB_GLOBAL(d.acc.mult_reads); SFB_GLOBAL(d.acc.mult_norm);
    SFB_GLOBAL(d.acc.mult_hits); SFB_GLOBAL(d.acc.uniq_abund); SFB_GLOBAL(d.acc.mult_abund); SFB_GLOBAL(d.acc.hist);
    SFB_GLOBAL(d.acc.counters);
    SFB_GLOBAL(t.tile_desc); SFB_GLOBAL(t.items); SFB_GLOBAL(t.edge_off); SFB_GLOBAL(t.edge_succ); SFB_GLOBAL(t.edge_mask);
    SFB_GLOBAL(t.node_mask); SFB_GLOBAL(U);
#undef SFB_GLOBAL
    TileS T = tile_carve(lay, t);
    BlockCounters& sc = *T.sc();
    block_counters_zero(sc);
    TileStats st{0, 0, 0, 0, 0, 0};
    float apg = 5.0f, dfrac = 0.3f;      // pass 2: running averages that size a round (block-uniform)
    const int4* items = reinterpret_cast<const int4*>(t.items);
    const int tid = threadIdx.x;
    for (;;) {
        __syncthreads();
        if (tid == 0) T.misc()[0] = (int)atomicAdd(d.w.cursor + (PASS == 1 ? 5 : 6), 1ull);
        __syncthreads();
        const int item = T.misc()[0];
        if (item >= (int)t.n_items) break;
        const int4 it = items[item];
        tile_stage<PASS>(T, d, t, it.x, U);
        T.a_item = d.r.grp_aln_off[2 * (i64)it.y];
        if (tid == 0) T.misc()[1] = 0;
        for (i64 g0 = it.y; g0 < (i64)it.z;) {
            if (tid < 4 && (PASS == 2 || tid > 0)) T.misc()[1 + tid] = 0;
            __syncthreads();
            if (PASS == 1) {
                const i64 g1 = tile_window(T, d.r, g0, it.z, kRows1, kTT, 2 * kTT);
                T.a_row0 = d.r.grp_aln_off[2 * g0];
                const int nA = (int)(d.r.grp_aln_off[2 * g1] - T.a_row0);
                if (nA > kRows1) {         // one group with more alignments than a window holds: not in tile order as staged
                    if (tid == 0) SFB_COUNT(sc, SFB_C_MATCH_OVERFLOW, 1);      // (sfb_tile_permute keeps such groups out)
                    g0 = g1;
                    continue;
                }
                // the window's walk offsets and records into shared memory (coalesced), aln_assign = "none"
                T.w_row0 = d.r.aln_walk_off[T.a_row0];
                for (int i = tid; i <= nA; i += kTT) T.woff()[i] = (uint32_t)(d.r.aln_walk_off[T.a_row0 + i] - T.w_row0);
                for (int i = tid; i < nA; i += kTT) d.w.aln_assign[T.a_row0 + i] = -1;
                {
                    const int nR = (int)min((i64)kStageRec, d.r.aln_walk_off[T.a_row0 + nA] - T.w_row0);
                    for (int i = tid; i < nR; i += kTT) T.wnode()[i] = (uint16_t)tile_local(T, d.r.walk_node[T.w_row0 + i]);
                }
                {   // the next window (about as large as this one) on its way to the L2 while this one is worked on
                    const i64 g2 = min(g1 + (g1 - g0), (i64)it.z);
                    const i64 a1 = T.a_row0 + nA, w1 = d.r.aln_walk_off[a1];
                    tile_prefetch(d.r, g1, g2, it.z);
                    tile_prefetch_alns(d.r, a1, min(a1 + nA, (i64)d.r.n_alns));
                    tile_prefetch_recs(d.r, w1, min(w1 + (w1 - T.w_row0), (i64)d.r.n_records));
                }
                __syncthreads();
                // M
                if (T.simple) for (int i = tid; i < nA; i += kTT) mask_match_aln<true>(T, d.r, i, T.a_row0 + i, st);
                else for (int i = tid; i < nA; i += kTT) tile_match_aln<KA>(T, d.r, i, T.a_row0 + i, st);
                __syncthreads();
                // C
                if (g0 + tid < g1) tile_group1<KA>(T, d, sc, g0 + tid, (int)(d.r.grp_aln_off[2 * (g0 + tid)] - T.a_row0), st);
                __syncthreads();
                // S: whole iterations only; the rest of the list stays for the next window (the last one takes all)
                g0 = g1;
                const int n_asg = T.misc()[1];
                const int full = g0 >= (i64)it.z ? n_asg : (n_asg / kTT) * kTT;
                for (int q = tid; q < full; q += kTT) tile_scatter1(T, d, sc, q, st);
                uint32_t ka = 0, kx = 0;
                int ks = 0;
                const bool keep = full + tid < n_asg;
                if (keep) { ka = T.la()[full + tid]; kx = T.lx()[full + tid]; ks = T.lseq()[full + tid]; }
                __syncthreads();
                if (keep) { T.la()[tid] = ka; T.lx()[tid] = kx; T.lseq()[tid] = ks; }
                if (tid == 0) T.misc()[1] = n_asg - full;
            } else {
                // D: windows of read groups until the deferred groups listed should fill phase S's list (apg = work
                //    items per deferred group, dfrac = share of deferred groups, running averages of this block)
                const int target = max(8, min(kTT, (int)(0.7f * (float)tile_list_cap(2) / apg)));
                for (;;) {
                    const int want = max(16, min(kTT, (int)((float)(target - T.misc()[3]) / fmaxf(dfrac, 0.03f))));
                    const i64 g1 = tile_window(T, d.r, g0, it.z, kWin2, want);
                    const i64 a0 = d.r.grp_aln_off[2 * g0];
                    const int before = T.misc()[3];
                    __syncthreads();       // everyone has read the fill
                    if (d.r.grp_aln_off[2 * g1] - a0 > kWin2) {                // see above
                        if (tid == 0) SFB_COUNT(sc, SFB_C_MATCH_OVERFLOW, 1);
                    } else if (g0 + tid < g1) {
                        const i64 g = g0 + tid;
                        const int c1 = d.w.grp_code[g];
                        if ((c1 & 0xf) == SFB_DEFER || (c1 >> 4) == SFB_DEFER) {
                            const i64 ga = d.r.grp_aln_off[2 * g];
                            const int na = (int)(d.r.grp_aln_off[2 * g + 2] - ga);
                            const int slot = atomicAdd(&T.misc()[3], 1), row = atomicAdd(&T.misc()[2], na);
                            T.glist()[slot] = (uint32_t)(g - it.y);
                            T.grow()[slot] = (uint16_t)row;
                            for (int j = 0; j < na; ++j) T.alist()[row + j] = (uint32_t)(ga + j - T.a_item);
                        }
                    }
                    __syncthreads();
                    dfrac = 0.5f * dfrac + 0.5f * (float)(T.misc()[3] - before) / (float)(g1 - g0);
                    g0 = g1;
                    if (g0 >= (i64)it.z || T.misc()[3] >= target - target / 8 || T.misc()[3] + kTT > kGrp2 || T.misc()[2] + kWin2 > kRows2)
                        break;
                }
                const int n_aln = T.misc()[2], n_grp = T.misc()[3];
                // M
                {
                    const int seen = st.tuples;                       // counted in pass 1
                    if (T.simple) for (int q = tid; q < n_aln; q += kTT) mask_match_aln<false>(T, d.r, q, T.a_item + T.alist()[q], st);
                    else for (int q = tid; q < n_aln; q += kTT) tile_match_aln<KA>(T, d.r, q, T.a_item + T.alist()[q], st);
                    st.tuples = seen;
                }
                __syncthreads();
                // C
                for (int q = tid; q < n_grp; q += kTT) tile_group2<KA>(T, d, sc, (i64)it.y + T.glist()[q], T.grow()[q], st);
                __syncthreads();
                // S: per-path sums (json2csv.py:1155-1156) and node loop (:1157-1158) of every work item
                const int n_app = min(T.misc()[1], tile_list_cap(2));
                if (n_grp) apg = 0.5f * apg + 0.5f * (float)T.misc()[1] / (float)n_grp;
                for (int q = tid; q < n_app; q += kTT) tile_scatter2(T, d, q, st);
            }
            __syncthreads();
        }
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
                    t->max_nodes >= 0 && t->max_nodes <= 65520 && t->max_edges >= 0 && t->max_edges <= 65535,
                "tile exceeds the limits (32767 slots, 255 paths, 65520 nodes, 65535 steps)");
    SFB_REQUIRE(t->n_items == 0 || (t->tile_desc && t->items && t->n_tiles > 0), "null tile array");
    SFB_REQUIRE(t->n_items == 0 || t->max_edges == 0 || (t->edge_off && t->edge_succ && t->edge_mask && t->node_mask),
                "null step index");
    return SFB_OK;
}

i64 tile_smem_bytes(const sfb_tiles* t) {
    if (tile_check(t) != SFB_OK) return SFB_E_ARG;
    const i64 a = tile_layout(1, (int)t->max_slots, (int)t->max_nodes, (int)t->max_paths, (int)t->max_edges, (int)t->tuple_cap).total;
    const i64 b = tile_layout(2, (int)t->max_slots, (int)t->max_nodes, (int)t->max_paths, (int)t->max_edges, (int)t->tuple_cap).total;
    return a > b ? a : b;
}

template <int PASS, int KA>
static int launch_tile_k(const Dev& d, const sfb_tiles* t, const long long* U, cudaStream_t st) {
    const TileLayout lay = tile_layout(PASS, (int)t->max_slots, (int)t->max_nodes, (int)t->max_paths, (int)t->max_edges, KA);
    const size_t smem = lay.total;
    auto kern = k_tile_pass<PASS, KA>;
    if (smem > 227 * 1024) return fail(SFB_E_ARG, "tile kernels: %zu bytes of shared memory needed, 232448 available", smem);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return check_launch("k_tile_pass (shared memory size)");
    int per_sm = 0, dev = 0, sms = kSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTT, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const i64 grid = t->n_items < (i64)per_sm * sms ? t->n_items : (i64)per_sm * sms;
    kern<<<(unsigned)grid, kTT, smem, st>>>(d, *t, lay, U);
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
