// k_match_steps: Pangenome.get_matching_path (json2csv.py:341-363) for the walk and the reversed walk through the step
// index of the simple tiles (sfb_graph.node_step in sfb200.h) instead of the occurrence index.
//
// One thread per alignment.  A path of a simple tile visits a node at most once, so the walk matches path p (at the
// position of its first node) iff p takes every step of the walk: the forward matches are the AND over the steps
// (n[k] -> n[k+1]) of the set of paths taking that step, the matches of the reversed walk the AND over the steps
// (n[k+1] -> n[k]).  Both sets come from one 32-byte record per walk node (its two commonest successors with their path
// sets): L record loads and 2(L-1) compares per alignment, whatever the number of haplotypes through the nodes, where
// the occurrence-index kernel tests every occurrence of the first and of the last node (10 candidates per alignment for 2
// matches on C3).  The two sets ARE the result (16 bytes per alignment): the group kernels work on them with AND / OR /
// popcount and turn a path into the reference's tuple -- (slot of the walk's first node on the path, path) for the walk,
// (slot of its last node | REV, path) for the reversed walk -- only where a read is assigned or shared (sets_tuple in
// common.cuh, used by group_sets.cuh; k_expand_sets below writes whole tuple lists for the few groups that take the generic kernels).
#include "common.cuh"

namespace sfb {

struct StepRec {
    int succ0, succ1;
    u64 mask0, mask1;
    int p0, n_succ;
};
__device__ __forceinline__ StepRec step_load(const sfb_graph& g, int n) {
    // the whole record with one 256-bit load (LDG.E.256 on sm_100): a gather costs the L1 one pass per distinct line
    // per instruction, so two 128-bit loads would cost twice
    const char* p = reinterpret_cast<const char*>(g.node_step) + 32 * (size_t)n;
    u64 q0, q1, q2, q3;
    asm("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(q0), "=l"(q1), "=l"(q2), "=l"(q3) : "l"(p));
    StepRec r;
    r.succ0 = (int)(uint32_t)q0;
    r.succ1 = (int)(uint32_t)(q0 >> 32);
    r.mask0 = q1;
    r.mask1 = q2;
    r.p0 = (int)(uint32_t)q3;
    r.n_succ = (int)(uint32_t)(q3 >> 32);
    return r;
}
// the paths that go from node `from` (record rc) straight to node `to`
__device__ __forceinline__ u64 step_mask(const sfb_graph& g, const StepRec& rc, int from, int to) {
    u64 m = 0;
    if (rc.succ0 == to) m = rc.mask0;
    if (rc.succ1 == to) m = rc.mask1;
    if (rc.n_succ > 2) {                                  // rare: more than two distinct successors
        const int e0 = g.step_off[from];
        for (int e = e0 + 2; e < e0 + rc.n_succ; ++e)
            if (g.step_succ[e] == to) m = g.step_mask[e];
    }
    return m;
}

constexpr int kStepBlock = 256;
#ifndef SFB_STEPS_BATCH
#define SFB_STEPS_BATCH 2
#endif
#ifndef SFB_STEPS_PREFETCH
#define SFB_STEPS_PREFETCH 1
#endif
#ifndef SFB_STEPS_MINBLK
#define SFB_STEPS_MINBLK 5
#endif
constexpr int kStepBatch = SFB_STEPS_BATCH ? SFB_STEPS_BATCH : 1;

__global__ void __launch_bounds__(kStepBlock, SFB_STEPS_MINBLK) k_match_steps(sfb_graph g, sfb_reads r, sfb_work w, long long* counters) {
    __shared__ BlockCounters sc;
    block_counters_zero(sc);
    __syncthreads();
    uint2* aln_match = reinterpret_cast<uint2*>(w.aln_match);
    ulonglong2* aln_sets = reinterpret_cast<ulonglong2*>(w.aln_sets);
    int32_t* fb_list = w.aln_assign + w.aln_begin;
    const i64 a_lo = w.aln_begin;
    // a block works through one contiguous stretch of alignments: its warps stay on neighbouring walks, hence (reads in
    // tile order) on the records of one part of the graph
    const i64 n_round = (r.n_alns - a_lo + kStepBlock - 1) / kStepBlock * kStepBlock;
    const i64 per = ((n_round / kStepBlock + gridDim.x - 1) / gridDim.x) * kStepBlock;
    const i64 b_lo = (i64)blockIdx.x * per, b_hi = min(b_lo + per, n_round);
    int st_steps = 0, st_tuples = 0;
    // one pass ahead: the first record of the walk this thread takes next (its walk offset is loaded a pass earlier), so
    // that the node ids of that walk are in the L1 when the pass begins
    i64 w_next = -1;
    if (SFB_STEPS_PREFETCH && a_lo + b_lo + threadIdx.x + kStepBlock < r.n_alns) w_next = r.aln_walk_off[a_lo + b_lo + threadIdx.x + kStepBlock];
    for (i64 base = b_lo + threadIdx.x; base < b_hi; base += kStepBlock) {
        const i64 a = a_lo + base;
        const bool live = a < r.n_alns;
        if (SFB_STEPS_PREFETCH) {
            if (w_next >= 0 && w_next < r.n_records) asm volatile("prefetch.global.L1 [%0];" ::"l"(r.walk_node + w_next));
            const i64 a2 = a + 2 * kStepBlock;
            w_next = (base + 2 * kStepBlock < b_hi && a2 < r.n_alns) ? r.aln_walk_off[a2] : -1;
        }
        u64 f = 0, rv = 0;
        int p0 = 0;
        bool fb = false;
        if (live) {
            const i64 w0 = r.aln_walk_off[a];
            const int L = (int)(r.aln_walk_off[a + 1] - w0);
            if (L > 0) {
                const int first = r.walk_node[w0];
                StepRec rec = step_load(g, first);
                p0 = rec.p0;
                if (p0 < 0) {
                    fb = true;
                } else if (L == 1) {
                    f = g.node_mask[first];
                } else {
                    f = rv = ~0ull;
                    int prev = first;
#if SFB_STEPS_BATCH
                    // kStepBatch nodes of the walk, then their records, in flight together: the walk costs two memory
                    // latencies per batch instead of two per node (a step is only counted while a set is alive, as
                    // the node-by-node loop with its early exit counts them)
                    for (int k0 = 1; k0 < L && (f | rv); k0 += kStepBatch) {
                        int n[kStepBatch];
                        StepRec rn[kStepBatch];
#pragma unroll
                        for (int j = 0; j < kStepBatch; ++j) n[j] = k0 + j < L ? r.walk_node[w0 + k0 + j] : -1;
#pragma unroll
                        for (int j = 0; j < kStepBatch; ++j)
                            if (n[j] >= 0) rn[j] = step_load(g, n[j]);
#pragma unroll
                        for (int j = 0; j < kStepBatch; ++j)
                            if (n[j] >= 0 && (f | rv)) {
                                f &= step_mask(g, rec, prev, n[j]);
                                rv &= step_mask(g, rn[j], n[j], prev);
                                st_steps += 1;
                                rec = rn[j];
                                prev = n[j];
                            }
                    }
#else
                    for (int k = 1; k < L; ++k) {
                        const int n = r.walk_node[w0 + k];
                        const StepRec rn = step_load(g, n);
                        f &= step_mask(g, rec, prev, n);
                        rv &= step_mask(g, rn, n, prev);
                        st_steps += 1;
                        if (!(f | rv)) break;
                        rec = rn;
                        prev = n;
                    }
#endif
                }
            }
            if (!fb) {
                if (f | rv) {
                    aln_sets[a] = make_ulonglong2(f, rv);
                    aln_match[a] = make_uint2(SFB_MATCH_SETS, (uint32_t)p0);
                    st_tuples += __popcll(f) + __popcll(rv);
                } else {
                    aln_match[a] = make_uint2(SFB_MATCH_NONE, 0u);
                }
            }
        }
        if (g.steps_partial) {                              // (kernel-uniform)
            const u64 fpos = warp_claim(w.cursor + 8, fb ? 1u : 0u);
            if (fb) fb_list[fpos] = (int32_t)a;
        }
    }
    const int c0 = warp_sum(st_steps), c1 = warp_sum(st_tuples);
    if (lane_id() == 0) {
        SFB_COUNT(sc, SFB_C_ST_COMPARES, c0);      // steps evaluated (both orientations from the same records)
        SFB_COUNT(sc, SFB_C_ST_TUPLES, c1);
    }
    __syncthreads();
    block_counters_flush(sc, counters);
}

// The path sets of the alignments of the read groups listed in w.slow[0 .. cursor[which]) (groups that take the generic
// kernel of classify, which = 3, or of pass 2, which = 4) written out as tuple lists, in the reference's order: forward
// matches by ascending path, then reverse.  One thread per group; rare.
__global__ void __launch_bounds__(128) k_expand_sets(sfb_graph g, sfb_reads r, sfb_work w, long long* counters, int which) {
    uint2* aln_match = reinterpret_cast<uint2*>(w.aln_match);
    uint2* match_buf = reinterpret_cast<uint2*>(w.match_buf);
    const i64 n_slow = (i64)w.cursor[which];
    for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < n_slow; q += (i64)gridDim.x * blockDim.x) {
        const i64 grp = w.slow[q];
        for (i64 a = r.grp_aln_off[2 * grp]; a < r.grp_aln_off[2 * grp + 2]; ++a) {
            const uint2 m = aln_match[a];
            if (m.x != SFB_MATCH_SETS) continue;
            const u64 f = w.aln_sets[2 * a], rv = w.aln_sets[2 * a + 1];
            const int n = __popcll(f) + __popcll(rv);
            const i64 w0 = r.aln_walk_off[a], w1 = r.aln_walk_off[a + 1];
            const int first = r.walk_node[w0], last = r.walk_node[w1 - 1];
            uint2 one = make_uint2(SFB_MATCH_NONE, 0u);
            uint2* out = &one;
            u64 pos = 0;
            if (n > 1) {
                pos = atomicAdd(w.cursor, (u64)n);
                if (pos + n > (u64)w.match_cap) {
                    aln_match[a] = make_uint2(SFB_MATCH_NONE, 0u);
                    atomicAdd((u64*)&counters[SFB_C_MATCH_OVERFLOW], (u64)n);
                    continue;
                }
                out = match_buf + pos;
            }
            for (u64 rest = f; rest; rest &= rest - 1) *out++ = sets_tuple(g, first, (int)m.y, rest, false);
            for (u64 rest = rv; rest; rest &= rest - 1) *out++ = sets_tuple(g, last, (int)m.y, rest, true);
            aln_match[a] = n > 1 ? make_uint2(SFB_MATCH_MULTI | (uint32_t)pos, (uint32_t)n) : one;
        }
    }
}

int launch_match_list(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, const int32_t* list,
                      const unsigned long long* n_list, cudaStream_t st);

int launch_match_steps(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, cudaStream_t st) {
    const i64 n = r->n_alns - w->aln_begin;
    if (n <= 0) return SFB_OK;
    static OccCache cache;
    const unsigned grid = persistent_grid(k_match_steps, kStepBlock, blocks_for(n, kStepBlock), cache);
    k_match_steps<<<grid, kStepBlock, 0, st>>>(*g, *r, *w, acc->counters);
    SFB_TRY(check_launch("k_match_steps"));
    if (!g->steps_partial) return SFB_OK;
    return launch_match_list(g, r, w, acc, w->aln_assign + w->aln_begin, w->cursor + 8, st);
}

int launch_expand_sets(const sfb_graph* g, const sfb_reads* r, const sfb_work* w, sfb_accum* acc, int which, cudaStream_t st) {
    static OccCache cache;
    const unsigned grid = persistent_grid(k_expand_sets, 128, blocks_for(r->n_groups, 128), cache);
    k_expand_sets<<<grid, 128, 0, st>>>(*g, *r, *w, acc->counters, which);
    return check_launch("k_expand_sets");
}

}  // namespace sfb
