// Node-level sums in shared memory for read groups staged in tile order, on top of the path-set kernels (steps.cu,
// group_sets.cuh): a thread block takes a work item (tile, range of read groups), keeps the tile's slot accumulators as
// zeroed 96-bit fixed-point bins in shared memory (bins.cuh), adds every record of the item there -- native 32-bit
// ATOMS instead of 64-bit REDs at the L2, whose rate bounds k_pass1_scatter / k_pass2_scatter -- and flushes the bins
// with one pair of integer REDs per touched slot.  Integer sums: the result does not depend on the order of anything.
//   k_tile_scatter1  fill_abund_dist's node loop (json2csv.py:329-330) for the unique assignments of pass 1
//   k_tile_share     pass 2 (json2csv.py:1085-1274) for the deferred groups of the item: candidate selection and ratios on
//                    the path sets (group_sets.cuh), per-path sums (:1155-1156) and node loops (:1157-1158) into the bins
// Nothing of the graph is staged: the few index gathers (slot of a path on a node) stay in L1.
#include <cstdlib>

#include "common.cuh"
#include "bins.cuh"
#include "group_sets.cuh"

namespace sfb {

constexpr int kBinBlock = 256;
#ifndef SFB_SCATTER1_TWO
#define SFB_SCATTER1_TWO 1
#endif

struct TileRange {
    int p0, np, s0, ns;
};
__device__ __forceinline__ TileRange tile_range(const sfb_graph& g, const sfb_tiles& t, int tile) {
    const int4 desc = reinterpret_cast<const int4*>(t.tile_desc)[2 * tile];
    TileRange tr;
    tr.p0 = desc.x;
    tr.np = desc.y - desc.x;
    tr.s0 = g.path_off[desc.x];
    tr.ns = g.path_off[desc.y] - tr.s0;
    return tr;
}

__global__ void __launch_bounds__(kBinBlock) k_tile_scatter1(Dev d, sfb_tiles t) {
    extern __shared__ __align__(16) uint32_t s_bins[];          // [3 * max_slots]
    __shared__ int s_item;
    const int4* items = reinterpret_cast<const int4*>(t.items);
    const int tid = threadIdx.x;
    int done = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = (int)atomicAdd(d.w.cursor + 5, 1ull);
        __syncthreads();
        const int item = s_item;
        if (item >= (int)t.n_items) break;
        const int4 it = items[item];
        const TileRange tr = tile_range(d.g, t, it.x);
        for (int i = tid; i < 3 * tr.ns; i += kBinBlock) s_bins[i] = 0u;
        __syncthreads();
        const i64 a0 = d.r.grp_aln_off[2 * (i64)it.y], a1 = d.r.grp_aln_off[2 * (i64)it.z];
#if SFB_SCATTER1_TWO
        // two alignments per thread and pass: their codes, then their walk offsets, then their first records are in flight
        // together (44 % of the alignments carry no assignment: one chain per thread leaves half the lanes without a load)
        for (i64 a = a0 + tid; a < a1; a += 2 * kBinBlock) {
            const i64 b = a + kBinBlock;
            const int32_t c0 = d.w.aln_assign[a], c1 = b < a1 ? d.w.aln_assign[b] : -1;
            i64 w0 = 0, w1 = 0;
            int L0 = 0, L1 = 0;
            if (c0 >= 0) {
                w0 = d.r.aln_walk_off[a];
                L0 = (int)(d.r.aln_walk_off[a + 1] - w0);
            }
            if (c1 >= 0) {
                w1 = d.r.aln_walk_off[b];
                L1 = (int)(d.r.aln_walk_off[b + 1] - w1);
            }
            int v0[4], v1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v0[u] = u < L0 ? d.r.walk_alen[w0 + u] : 0;
                v1[u] = u < L1 ? d.r.walk_alen[w1 + u] : 0;
            }
            Fx one;
            one.x0 = one.x1 = 0u;
            one.x2 = 1u;
            uint32_t* b0 = s_bins + 3 * (((uint32_t)c0 & SFB_SLOT_MASK) - tr.s0);
            uint32_t* b1 = s_bins + 3 * (((uint32_t)c1 & SFB_SLOT_MASK) - tr.s0);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (u < L0) bins_add_record(b0 + 3 * u, d.r, v0[u], true, 1.0, one);
                if (u < L1) bins_add_record(b1 + 3 * u, d.r, v1[u], true, 1.0, one);
            }
            if (L0 > 4) bins_add_records(b0 + 12, d.r, w0 + 4, L0 - 4, true, 1.0);
            if (L1 > 4) bins_add_records(b1 + 12, d.r, w1 + 4, L1 - 4, true, 1.0);
            done += L0 + L1;
        }
#else
        for (i64 a = a0 + tid; a < a1; a += kBinBlock) {
            const int32_t code = d.w.aln_assign[a];
            if (code < 0) continue;
            const i64 w0 = d.r.aln_walk_off[a];
            const int L = (int)(d.r.aln_walk_off[a + 1] - w0);
            bins_add_records(s_bins + 3 * (((uint32_t)code & SFB_SLOT_MASK) - tr.s0), d.r, w0, L, true, 1.0);
            done += L;
        }
#endif
        __syncthreads();
        for (int i = tid; i < tr.ns; i += kBinBlock) fx_flush(d.acc.uniq_abund + tr.s0 + i, d.acc.det_stride, s_bins + 3 * i);
    }
    done = warp_sum(done);
    if (lane_id() == 0 && done) atomicAdd((u64*)&d.acc.counters[SFB_C_ST_P1_RECORDS], (u64)done);
}

// ---- pass 2 ---------------------------------------------------------------------------------------------------------
// Per warp, three kinds of steps, each with one dense unit of work per lane (a thread per deferred group doing its own
// tuple and record loops ran at 6 active lanes of 32):
//   scan    32 read groups of the item: the deferred ones join the warp's queue
//   decide  32 queued groups, lane = group: candidate selection and ratios on the path sets; every share becomes an entry
//           of the warp's share list (alignment, path, orientation, ratio, times, len(sequence))
//   add     32 listed shares, lane = share: per-path sums, slot of the path on the walk's first / last node; their node
//           loops are then laid over the warp (share_records)
#ifndef SFB_SHARE_BLOCK
#define SFB_SHARE_BLOCK 256                 /* measured on C3 (profiles/README.md r02r-r02t): 8 warps share one set of bins, 4 blocks
                                               per SM = 32 warps at 64 registers (some spills) beat 20-24 warps at 80-96 */
#endif
constexpr int kShareBlock = SFB_SHARE_BLOCK;
constexpr int kShareWarps = kShareBlock / 32;
constexpr int kQueue = 64;                 // deferred groups waiting per warp (a scan adds at most 32)
#ifndef SFB_SHARE_FLAT
#define SFB_SHARE_FLAT 1
#endif
#ifndef SFB_SHARE_LIST
#define SFB_SHARE_LIST 160
#endif
constexpr int kShares = SFB_SHARE_LIST;    // entries of a warp's share list (a round of 32 groups lists ~160 shares on C3; 256
                                           // entries cost a block per SM, 128 / 96 are slower: 3.56 / 3.63 against 3.49 ms)
constexpr int kDirect = kShares;           // a group with more shares than this adds them itself (no list)
constexpr uint32_t kShareHist = 0x80000000u;

// What a share needs beyond its path: the candidate list it belongs to.  One plan per (group of the round, mate,
// alignment); the ratio, the multiplicity and the histogram update of a share are computed by the lane that adds it.
struct SharePlan {
    long long total;      // sum of the unique counts over the candidate list
    u64 common;           // != 0: strain-reduced list, a tuple counts popcount(smask[path] & common) times (Q4)
    int n;                // length of the list
    int seq_len;
};
struct ShareList {
    SharePlan plan[128];
    uint32_t a_off[kShares];               // alignment, from the first alignment of the work item
    uint32_t path[kShares];                // path id | SFB_MATCH_REV | kShareHist
    uint16_t plan_of[kShares];
};
// one share into the bins: per-path sums (json2csv.py:1155-1156; three limbs each + "touched") and the node loop (:1157-1158)
struct BinEmit {
    const Dev& d;
    uint32_t* bins;
    uint32_t* plimb;      // [np][7]: touched, mult_reads limbs, mult_norm limbs
    TileRange tr;
    int rec, app;
    __device__ __forceinline__ void operator()(i64 a, bool rev, int pid, int seq_len, double ratio, int times) {
        const double weight = ratio * (double)times;
        uint32_t* l = plimb + 7 * (pid - tr.p0);
        l[0] = 1u;
        fx_smem_add(l + 1, fx_split(weight));
        fx_smem_add(l + 4, fx_split((ratio * (double)seq_len / (double)d.g.path_seq_len[pid]) * (double)times));
        const i64 w0 = d.r.aln_walk_off[a];
        const int L = (int)(d.r.aln_walk_off[a + 1] - w0);
        const int node = d.r.walk_node[rev ? w0 + L - 1 : w0];
        const uint2 tup = sets_tuple(d.g, node, tr.p0, 1ull << (pid - tr.p0), rev);
        bins_add_records(bins + 3 * ((tup.x & SFB_SLOT_MASK) - tr.s0), d.r, w0, L, false, weight);
        rec += L;
        app += 1;
    }
    // the same without the node loop: where the share's records start (bin offset in words, first record, count) and its
    // weight, for share_records below
    __device__ __forceinline__ void share(i64 a, bool rev, int pid, int seq_len, double ratio, int times, uint32_t& bo, i64& w0,
                                          int& L, double& weight) {
        weight = ratio * (double)times;
        uint32_t* l = plimb + 7 * (pid - tr.p0);
        l[0] = 1u;
        fx_smem_add(l + 1, fx_split(weight));
        fx_smem_add(l + 4, fx_split((ratio * (double)seq_len / (double)d.g.path_seq_len[pid]) * (double)times));
        w0 = d.r.aln_walk_off[a];
        L = (int)(d.r.aln_walk_off[a + 1] - w0);
        const int node = d.r.walk_node[rev ? w0 + L - 1 : w0];
        const uint2 tup = sets_tuple(d.g, node, tr.p0, 1ull << (pid - tr.p0), rev);
        bo = 3u * ((tup.x & SFB_SLOT_MASK) - (uint32_t)tr.s0);
        rec += L;
        app += 1;
    }
};
// The node loops (json2csv.py:1157-1158) of the 32 shares a warp holds, one per lane (L = 0: none).  A lane that walked its
// own records ran beside lanes with shorter walks (15 of 32 lanes busy, 45 % of the kernel's instructions): here the first
// and the last record of every share -- the two that usually cover their node only in part, cov * weight with an fp64
// division -- are taken lane = share, and the records in between, which add the share's weight as it is, are laid end to
// end over the warp: lane = record, its share found by a binary search in the running sum of the interior counts
// (shuffles), every lane busy whatever the walk lengths.  Called by the whole warp.
__device__ __forceinline__ void share_records(uint32_t* bins, const sfb_reads& r, uint32_t bo, i64 w0, int L, double weight) {
    const Fx whole = fx_split(weight);
    int v0 = 0, v1 = 0;
    if (L > 0) v0 = r.walk_alen[w0];
    if (L > 1) v1 = r.walk_alen[w0 + L - 1];
    if (L > 0) bins_add_record(bins + bo, r, v0, false, weight, whole);
    if (L > 1) bins_add_record(bins + bo + 3 * (L - 1), r, v1, false, weight, whole);
    const int inner = L > 2 ? L - 2 : 0;
    const int incl = (int)warp_scan((unsigned)inner);
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - inner;
    const uint32_t w_lo = (uint32_t)(u64)w0, w_hi = (uint32_t)((u64)w0 >> 32);
    for (int r0 = 0; r0 < total; r0 += 32) {
        const int q = r0 + lane_id();
        int s = 0;                               // the lanes whose running sum is <= q: q belongs to the next one
#pragma unroll
        for (int step = 16; step > 0; step >>= 1)
            if (__shfl_sync(0xffffffffu, incl, s + step - 1) <= q) s += step;
        const int i = 1 + q - __shfl_sync(0xffffffffu, excl, s);
        const i64 ws = (i64)(((u64)__shfl_sync(0xffffffffu, w_hi, s) << 32) | __shfl_sync(0xffffffffu, w_lo, s));
        const uint32_t bs = __shfl_sync(0xffffffffu, bo, s);
        Fx f;
        f.x0 = __shfl_sync(0xffffffffu, whole.x0, s);
        f.x1 = __shfl_sync(0xffffffffu, whole.x1, s);
        f.x2 = __shfl_sync(0xffffffffu, whole.x2, s);
        const double wt = __shfl_sync(0xffffffffu, weight, s);
        if (q < total) bins_add_record(bins + bs + 3 * i, r, r.walk_alen[ws + i], false, wt, f);
    }
}

#ifndef SFB_SHARE_MINBLK
#define SFB_SHARE_MINBLK 4
#endif
__global__ void __launch_bounds__(kShareBlock, SFB_SHARE_MINBLK) k_tile_share(Dev d, sfb_tiles t, const long long* __restrict__ U) {
    extern __shared__ __align__(16) uint32_t s_mem[];           // bins [3 * max_slots] | plimb [7 * max_paths] | share lists
    __shared__ BlockCounters sc;
    __shared__ int s_item, s_chunk;
    __shared__ int s_queue[kShareWarps][kQueue];
    uint32_t* s_bins = s_mem;
    uint32_t* s_plimb = s_mem + 3 * (int)t.max_slots;
    const int tid = threadIdx.x, wi = tid >> 5, lane = lane_id();
    ShareList& list = reinterpret_cast<ShareList*>(s_mem + ((3 * (int)t.max_slots + 7 * (int)t.max_paths + 3) & ~3))[wi];
    block_counters_zero(sc);
    const int4* items = reinterpret_cast<const int4*>(t.items);
    const unsigned below = (1u << lane) - 1u;
    int st_rec = 0, st_app = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_item = (int)atomicAdd(d.w.cursor + 6, 1ull);
            s_chunk = 0;
        }
        __syncthreads();
        const int item = s_item;
        if (item >= (int)t.n_items) break;
        const int4 it = items[item];
        const TileRange tr = tile_range(d.g, t, it.x);
        for (int i = tid; i < 3 * tr.ns; i += kShareBlock) s_bins[i] = 0u;
        for (int i = tid; i < 7 * tr.np; i += kShareBlock) s_plimb[i] = 0u;
        __syncthreads();
        const i64 a_item = d.r.grp_aln_off[2 * (i64)it.y];
        BinEmit add{d, s_bins, s_plimb, tr, 0, 0};
        const int n_chunks = (it.z - it.y + 31) / 32;
        int qn = 0, fill = 0;                                   // warp-uniform fills of the queue and of the share list
        // add: the listed shares, 32 at a time, one per lane
        auto drain = [&]() {
            __syncwarp();
            for (int e0 = 0; e0 < fill; e0 += 32) {
                const int e = e0 + lane;
#if SFB_SHARE_FLAT
                uint32_t bo = 0;
                i64 w0 = 0;
                int L = 0;
                double weight = 0.0;
#endif
                if (e < fill) {
                    const uint32_t pth = list.path[e];
                    const SharePlan pl = list.plan[list.plan_of[e]];
                    const int pid = (int)(pth & SFB_SLOT_MASK);
                    const i64 a = a_item + list.a_off[e];
                    if (pth & kShareHist) hist_add(d, sc, pid, a);                                  // json2csv.py:1246-1252
                    const int times = pl.common ? __popcll(d.g.path_smask[pid] & pl.common) : 1;
#if SFB_SHARE_FLAT
                    add.share(a, (pth & SFB_MATCH_REV) != 0, pid, pl.seq_len, ratio_of(U[pid], pl.total, pl.n), times, bo, w0, L, weight);
#else
                    add(a, (pth & SFB_MATCH_REV) != 0, pid, pl.seq_len, ratio_of(U[pid], pl.total, pl.n), times);
#endif
                }
#if SFB_SHARE_FLAT
                __syncwarp();
                share_records(s_bins, d.r, bo, w0, L, weight);
#endif
            }
            fill = 0;
            __syncwarp();
        };
        for (bool last = false; !last;) {
            // scan
            int chunk = 0;
            if (lane == 0) chunk = atomicAdd(&s_chunk, 1);
            chunk = __shfl_sync(0xffffffffu, chunk, 0);
            last = chunk >= n_chunks;
            if (!last) {
                const i64 g = (i64)it.y + 32 * (i64)chunk + lane;
                bool defer = false;
                if (g < (i64)it.z) {
                    const int c1 = d.w.grp_code[g];
                    defer = (c1 & 0xf) == SFB_DEFER || (c1 >> 4) == SFB_DEFER;
                }
                const unsigned ball = __ballot_sync(0xffffffffu, defer);
                if (defer) {
                    s_queue[wi][qn + __popc(ball & below)] = (int)(g - (i64)it.y);
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(d.r.grp_aln_off + 2 * g));      // its header, for the decide step
                }
                qn += __popc(ball);
                __syncwarp();
            }
            // decide
            while (qn >= 32 || (last && qn > 0)) {
                const int take = min(qn, 32);
                qn -= take;
                Mate mt[2];
                SetMate sm[2];
                SetPlan sp;
                sp.mode[0] = sp.mode[1] = kNone;
                sp.sel[0] = sp.sel[1] = 0;
                sp.common = 0;
                int need = 0;
                if (lane < take) {
                    const i64 g = (i64)it.y + s_queue[wi][qn + lane];
                    const int nm = d.r.grp_nmates[g];
                    mate_load(d.r, g, 0, mt[0]);
                    mate_load(d.r, g, 1, mt[1]);
                    if (d.g.mask_words == 1 && sets_load(d.w, mt[0], sm[0]) && sets_load(d.w, mt[1], sm[1])) {
                        int code[2] = {SFB_P2_NONE, SFB_P2_NONE};
                        int single = nm == 1 ? 0 : -1;
                        if (nm == 2) {
                            if (mt[0].na == 0 || mt[1].na == 0) single = mt[0].na == 0 ? 1 : 0;          // json2csv.py:1088-1093
                            else if (sm[0].nt == 0 || sm[1].nt == 0) single = sm[0].nt == 0 ? 1 : 0;    // :1117-1122
                            else sets_p2_decide_pair(d, sc, sm, U, sp, code);
                        }
                        if (single == 0) {
                            sp.mode[0] = kSingle;
                            code[0] = p2_single_outcome(sc, sm[0].nt);
                        } else if (single == 1) {
                            sp.mode[1] = kSingle;
                            code[1] = p2_single_outcome(sc, sm[1].nt);
                        }
                        d.w.grp_code2[g] = (uint8_t)(code[0] | (code[1] << 4));
                        need = sets_p2_need(sm[0], sp.mode[0], sp.sel[0]) + sets_p2_need(sm[1], sp.mode[1], sp.sel[1]);
                        if (need) {       // the walk offsets of the group's alignments on their way to the L1 for the add steps
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(d.r.aln_walk_off + mt[0].a0));
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(d.r.aln_walk_off + mt[1].a0 + mt[1].na));
                        }
                        if (need > kDirect) {                    // (rare) more shares than a list takes: added here and now
                            sets_p2_emit(d, sc, mt[0], sm[0], sp, 0, U, add);
                            sets_p2_emit(d, sc, mt[1], sm[1], sp, 1, U, add);
                            need = 0;
                        }
                    } else {
                        // tuple lists (the group took the generic classify kernel): left to sfb_pass2_resolve
                        d.w.deferred[atomicAdd(d.w.cursor + 1, 1ull)] = (int32_t)g;
                        atomicAdd(d.w.cursor + 7, ~0ull);
                    }
                }
                if (need > 0) {
                    // the candidate lists of this lane's group: plan 4 * lane + 2 * mate + alignment
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        if (sp.mode[k] == kNone) continue;
                        SharePlan p0, p1;
                        p0.seq_len = p1.seq_len = mt[k].seq_len;
                        if (sp.mode[k] == kSingle) {
                            sets_single_totals(sm[k], U, p0.total, p1.total);
                            p0.n = __popcll(sm[k].m[0]) + __popcll(sm[k].m[1]);
                            p1.n = __popcll(sm[k].m[2]) + __popcll(sm[k].m[3]);
                            p0.common = p1.common = 0;
                        } else {
                            p0.total = p1.total = sp.total[k];
                            p0.n = p1.n = sp.n[k];
                            p0.common = p1.common = sp.mode[k] == kStrain ? sp.common : 0ull;
                        }
                        list.plan[4 * lane + 2 * k] = p0;
                        list.plan[4 * lane + 2 * k + 1] = p1;
                    }
                }
                // list the shares: the lanes whose shares fit (a prefix of the lanes), add what is listed, go on
                bool pending = need > 0;
                while (__any_sync(0xffffffffu, pending)) {
                    const unsigned incl = warp_scan(pending ? (unsigned)need : 0u);
                    int at = fill + (int)incl - need;
                    const bool fits = pending && at + need <= kShares;
                    if (fits) {
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            if (sp.mode[k] == kNone) continue;
                            const uint32_t flag = sp.mode[k] == kSingle ? kShareHist : 0u;
                            const i64 a0 = mt[k].a0 - a_item;
                            sets_for_tuples(sm[k], sp.mode[k] == kSingle ? ~0ull : sp.sel[k], [&](int q, int pid) {
                                list.a_off[at] = (uint32_t)(a0 + (q >> 1));
                                list.path[at] = (uint32_t)pid | ((q & 1) ? SFB_MATCH_REV : 0u) | flag;
                                list.plan_of[at] = (uint16_t)(4 * lane + 2 * k + (q >> 1));
                                ++at;
                            });
                        }
                        pending = false;
                    }
                    const unsigned okm = __ballot_sync(0xffffffffu, fits);
                    const int top = okm ? 31 - __clz((int)okm) : -1;
                    fill += top >= 0 ? (int)__shfl_sync(0xffffffffu, incl, top) : 0;
                    drain();                                     // (the plans belong to this round: nothing stays listed)
                }
            }
        }
        st_rec += add.rec;
        st_app += add.app;
        __syncthreads();
        for (int i = tid; i < tr.ns; i += kShareBlock) fx_flush(d.acc.mult_abund + tr.s0 + i, d.acc.det_stride, s_bins + 3 * i);
        for (int p = tid; p < tr.np; p += kShareBlock) {
            const uint32_t* l = s_plimb + 7 * p;
            if (!l[0]) continue;
            d.acc.mult_hits[tr.p0 + p] = 1u;
            fx_flush(d.acc.mult_reads + tr.p0 + p, d.acc.det_stride, l + 1);
            fx_flush(d.acc.mult_norm + tr.p0 + p, d.acc.det_stride, l + 4);
        }
    }
    const int c0 = warp_sum(st_rec), c1 = warp_sum(st_app);
    if (lane == 0) {
        SFB_COUNT(sc, SFB_C_ST_P2_RECORDS, c0);
        SFB_COUNT(sc, SFB_C_ST_P2_APPLIED, c1);
    }
    __syncthreads();
    block_counters_flush(sc, d.acc.counters);
}

static int bins_check(const sfb_tiles* t) {
    SFB_REQUIRE(t, "tiles is null");
    SFB_REQUIRE(t->n_tiles >= 0 && t->n_items >= 0 && t->n_items < 0x7fffffffLL, "bad tile / item count");
    SFB_REQUIRE(t->max_slots >= 0 && t->max_slots <= 32767 && t->max_paths >= 0 && t->max_paths <= 255, "tile exceeds the limits");
    SFB_REQUIRE(t->n_items == 0 || (t->tile_desc && t->items && t->n_tiles > 0), "null tile array");
    return SFB_OK;
}
template <class K>
static unsigned bins_grid(K kern, int block, size_t smem, i64 n_items, int at_most = 1 << 30) {
    int per_sm = 0, dev = 0, sms = kSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (per_sm > at_most) per_sm = at_most;
    return (unsigned)(n_items < (i64)per_sm * sms ? n_items : (i64)per_sm * sms);
}

int launch_tile_scatter1(const Dev& d, const sfb_tiles* t, cudaStream_t st) {
    SFB_TRY(bins_check(t));
    if (t->n_items == 0) return SFB_OK;
    const size_t smem = 12 * (size_t)t->max_slots;
    if (cudaFuncSetAttribute(k_tile_scatter1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return check_launch("k_tile_scatter1 (shared memory size)");
    // (measured: capping the blocks per SM so that the kernel could run beside sfb_tile_share on another stream does not
    // make them overlap -- it then simply runs longer afterwards; SFB200_SCATTER1_BLOCKS repeats the experiment)
    const char* env = getenv("SFB200_SCATTER1_BLOCKS");
    const int at_most = env ? atoi(env) : 0;
    k_tile_scatter1<<<bins_grid(k_tile_scatter1, kBinBlock, smem, t->n_items, at_most > 0 ? at_most : 1 << 30), kBinBlock, smem, st>>>(d, *t);
    return check_launch("k_tile_scatter1");
}
int launch_tile_share(const Dev& d, const sfb_tiles* t, const long long* U, cudaStream_t st) {
    SFB_TRY(bins_check(t));
    if (t->n_items == 0) return SFB_OK;
    const size_t smem = ((12 * (size_t)t->max_slots + 28 * (size_t)t->max_paths + 15) & ~(size_t)15) + kShareWarps * sizeof(ShareList);
    if (cudaFuncSetAttribute(k_tile_share, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return check_launch("k_tile_share (shared memory size)");
    k_tile_share<<<bins_grid(k_tile_share, kShareBlock, smem, t->n_items), kShareBlock, smem, st>>>(d, *t, U);
    return check_launch("k_tile_share");
}

}  // namespace sfb
