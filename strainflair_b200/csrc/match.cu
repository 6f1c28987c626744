// k_match: Pangenome.get_matching_path (json2csv.py:341-363) for the walk and, if longer than one node,
// the reversed walk (callers :855-866, :1024-1026, :1103-1114, :1231-1233).
//
// The node -> (slot, successor) occurrence index replaces Node.traversed_path and the linear position scan; a
// candidate is rejected by the successor stored in the index or at its first differing node, and the sentinel
// slot after every path rejects walks that run off the path end.  The walk arrays stream from HBM once; the index
// is gathered (partly L2 resident).  Output: 8 bytes per alignment, plus one (code, path id) pair per match for
// alignments with several matches, in space claimed with one atomic per warp.
#include "common.cuh"

namespace sfb {

#ifndef SFB_MATCH_KEEP
#define SFB_MATCH_KEEP 10     /* measured: 9-10 best on C3 (no list of its 2-9 haplotypes falls to phase C); more shared memory costs L1 */
#endif
constexpr int kKeep = SFB_MATCH_KEEP;   // matches staged per alignment; longer lists are produced by a second sweep

struct MatchStats {
    int cand, cmp;
};

// Candidate-parallel kernel, two phases per warp pass (the kernel is bound by memory latency, so the structure
// minimises dependent round trips and keeps lanes dense):
//   A  A warp takes 32 consecutive alignments (coalesced loads of their offsets and end nodes) and lays all their
//      candidate occurrences -- forward ones first, then reverse, per alignment -- on one index line; it walks that
//      line 64 candidates at a time, one lane = one candidate whatever alignment it belongs to, loads the
//      (slot, successor) pairs and keeps the candidates whose successor equals the walk's second node.  Survivors
//      are compacted (ballot) into a shared-memory list, in candidate order.
//   B  The list is drained 32 survivors at a time, all lanes busy: the remaining walk positions are fetched
//      together and compared, matches get their path id and their rank inside their alignment (ballot +
//      __match_any over the owner lane; candidate order = the reference's match order) and are staged in shared
//      memory, kKeep per alignment.
//   C  Alignments with more than kKeep matches are redone by the whole warp, one at a time, one lane per candidate,
//      and written straight to their place in match_buf.
#ifndef SFB_MATCH_KU
#define SFB_MATCH_KU 4
#endif
#ifndef SFB_MATCH_WIN
#define SFB_MATCH_WIN 2
#endif
#ifndef SFB_MATCH_MINBLK
#define SFB_MATCH_MINBLK 5      /* 48 registers: measured best occupancy / spill trade-off (profiles/README.md) */
#endif
constexpr int kWarpsPerBlock = 8;
constexpr int kU = SFB_MATCH_KU;    // compare positions fetched together (walks of up to kU + 2 nodes need no loop)
constexpr int kWin = SFB_MATCH_WIN; // windows of 32 candidates per phase-A trip
#ifndef SFB_MATCH_SURV_EXTRA
#define SFB_MATCH_SURV_EXTRA 96
#endif
constexpr int kSurv = 32 * kWin + SFB_MATCH_SURV_EXTRA;   // survivor list entries per warp

// `list` (optional): the alignments to match, *n_list of them (left over by k_match_steps); else all from w.aln_begin on.
__global__ void __launch_bounds__(256, SFB_MATCH_MINBLK) k_match(sfb_graph g, sfb_reads r, sfb_work w, long long* counters,
                                                                 const int32_t* __restrict__ list,
                                                                 const unsigned long long* __restrict__ n_list) {
    __shared__ BlockCounters sc;
    __shared__ uint2 s_stage[kWarpsPerBlock][32][kKeep];
    __shared__ uint2 s_surv[kWarpsPerBlock][kSurv];
    __shared__ int s_cnt[kWarpsPerBlock][32];
    block_counters_zero(sc);
    __syncthreads();
    MatchStats st{0, 0};
    int tuples = 0;
    uint2* aln_match = reinterpret_cast<uint2*>(w.aln_match);
    uint2* match_buf = reinterpret_cast<uint2*>(w.match_buf);
    const int2* occ = reinterpret_cast<const int2*>(g.occ_pair);
    const int wi = threadIdx.x >> 5, lane = lane_id();
    const unsigned below = (1u << lane) - 1u;
    const int last_slot = (int)g.n_slots - 1;
    const i64 a_lo = w.aln_begin;                       // the alignments before it belong to the tile kernels
    const i64 n_todo = list ? (i64)*n_list : r.n_alns - a_lo;
    const i64 n_round = (n_todo + 31) & ~(i64)31;
    for (i64 base = ((i64)blockIdx.x * kWarpsPerBlock + wi) * 32; base < n_round;
         base += (i64)gridDim.x * kWarpsPerBlock * 32) {
        const bool live = base + lane < n_todo;
        const i64 a = !live ? 0 : list ? (i64)list[base + lane] : a_lo + base + lane;
        i64 w0 = 0;
        int L = 0, second = 0, penult = 0, fo = 0, fcnt = 0, ro = 0, rcnt = 0;
        if (live) {
            w0 = r.aln_walk_off[a];
            L = (int)(r.aln_walk_off[a + 1] - w0);
            if (L > 0) {
                const int first = r.walk_node[w0];
                fo = g.occ_off[first];
                fcnt = g.occ_off[first + 1] - fo;
                if (L > 1) {
                    const int last = r.walk_node[w0 + L - 1];
                    second = r.walk_node[w0 + 1];
                    penult = r.walk_node[w0 + L - 2];
                    ro = g.occ_off[last];
                    rcnt = g.occ_off[last + 1] - ro;
                }
            }
        }
        const unsigned cnt = (unsigned)(fcnt + rcnt);
        const unsigned incl = warp_scan(cnt);
        const unsigned excl = incl - cnt;
        const int total = (int)__shfl_sync(0xffffffffu, incl, 31);
        s_cnt[wi][lane] = 0;
        __syncwarp();
        int t = 0, nsurv = 0;
        while (t < total) {
            // ---- phase A: successor filter, two windows of 32 candidates per trip ----
            while (t < total && nsurv <= kSurv - 32 * kWin) {
                int2 e[kWin];
                int own[kWin];
                bool rv[kWin], in[kWin];
                int succ[kWin], oL[kWin];
#pragma unroll
                for (int u = 0; u < kWin; ++u) {
                    const int c = t + 32 * u + lane;
                    int j = 0;                 // owner of candidate c: the largest lane j with excl_j <= c
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1) {
                        const unsigned x = __shfl_sync(0xffffffffu, excl, (j + step) & 31);
                        if (j + step < 32 && (int)x <= c) j += step;
                    }
                    const int k = c - (int)__shfl_sync(0xffffffffu, excl, j);
                    const int o_fcnt = __shfl_sync(0xffffffffu, fcnt, j);
                    const int o_fo = __shfl_sync(0xffffffffu, fo, j);
                    const int o_ro = __shfl_sync(0xffffffffu, ro, j);
                    const int o_second = __shfl_sync(0xffffffffu, second, j);
                    const int o_penult = __shfl_sync(0xffffffffu, penult, j);
                    oL[u] = __shfl_sync(0xffffffffu, L, j);
                    own[u] = j;
                    rv[u] = k >= o_fcnt;
                    in[u] = c < total;
                    succ[u] = rv[u] ? o_penult : o_second;
                    e[u] = in[u] ? occ[rv[u] ? o_ro + (k - o_fcnt) : o_fo + k] : make_int2(0, 0);
                }
#pragma unroll
                for (int u = 0; u < kWin; ++u) {
                    const bool keep = in[u] && (oL[u] == 1 || e[u].y == succ[u]);   // else rejected from the index
                    if (in[u]) { st.cand += 1; st.cmp += oL[u] > 1; }
                    const unsigned ball = __ballot_sync(0xffffffffu, keep);
                    if (keep)
                        s_surv[wi][nsurv + __popc(ball & below)] =
                            make_uint2((uint32_t)e[u].x | (rv[u] ? SFB_MATCH_REV : 0u), (uint32_t)own[u]);
                    nsurv += __popc(ball);
                }
                t += 32 * kWin;
            }
            __syncwarp();
            // ---- phase B: full compare of the survivors, 32 at a time ----
            for (int b = 0; b < nsurv; b += 32) {
                const bool valid = b + lane < nsurv;
                const uint2 sv = valid ? s_surv[wi][b + lane] : make_uint2(0u, 0u);
                const int j = (int)sv.y;
                const int slot = (int)(sv.x & SFB_SLOT_MASK);
                const bool rev = (sv.x & SFB_MATCH_REV) != 0;
                const int o_L = __shfl_sync(0xffffffffu, L, j);
                const i64 o_w0 = __shfl_sync(0xffffffffu, w0, j);
                bool matched = false;
                int pid = 0;
                int2 blk = make_int2(0, 0);
                if (valid) {
                    const i64 wf = rev ? o_w0 + o_L - 1 : o_w0;
                    const int step = rev ? -1 : 1;
                    blk = reinterpret_cast<const int2*>(g.blk_path)[slot >> 6];   // issued with the compare loads
                    int pn[kU], wn[kU];
#pragma unroll
                    for (int u = 0; u < kU; ++u) {
                        const int i = u + 2;
                        pn[u] = wn[u] = 0;
                        if (i < o_L) {
                            pn[u] = g.slot_node[min(slot + i, last_slot)];   // sentinel (-1) stops overruns
                            wn[u] = r.walk_node[wf + (i64)step * i];
                        }
                    }
                    int bad = o_L;                   // first differing position
#pragma unroll
                    for (int u = kU - 1; u >= 0; --u)
                        if (u + 2 < o_L && pn[u] != wn[u]) bad = u + 2;
                    for (int i = kU + 2; i < o_L && bad == o_L; ++i)                 // walks longer than kU + 2 nodes
                        if (g.slot_node[min(slot + i, last_slot)] != r.walk_node[wf + (i64)step * i]) bad = i;
                    if (o_L > 2) st.cmp += min(bad, o_L - 1) - 1;
                    matched = bad >= o_L;
                    if (matched) pid = path_of_block(g, blk, slot);
                }
                const unsigned ball = __ballot_sync(0xffffffffu, matched);
                const unsigned same = __match_any_sync(0xffffffffu, valid ? j : 32 + lane);
                const int before = valid ? s_cnt[wi][j] : 0;
                const int rank = before + __popc(ball & same & below);
                if (matched && rank < kKeep) s_stage[wi][j][rank] = make_uint2(sv.x, (uint32_t)pid);
                __syncwarp();
                if (matched && (ball & same & ~below & ~(1u << lane)) == 0) s_cnt[wi][j] = rank + 1;   // last match of its alignment here
                __syncwarp();
            }
            nsurv = 0;
        }
        __syncwarp();
        const int n = s_cnt[wi][lane];
        tuples += n;
        const unsigned need = n > 1 ? (unsigned)n : 0u;
        const u64 pos = warp_claim(w.cursor, need);
        const bool fits = pos + need <= (u64)w.match_cap;
        if (live) {
            if (n == 0) {
                aln_match[a] = make_uint2(SFB_MATCH_NONE, 0u);
            } else if (n == 1) {
                aln_match[a] = s_stage[wi][lane][0];
            } else if (fits) {
                if (n <= kKeep)
                    for (int q = 0; q < n; ++q) match_buf[pos + q] = s_stage[wi][lane][q];
                aln_match[a] = make_uint2(SFB_MATCH_MULTI | (uint32_t)pos, (uint32_t)n);
            } else {
                aln_match[a] = make_uint2(SFB_MATCH_NONE, 0u);
                SFB_COUNT(sc, SFB_C_MATCH_OVERFLOW, (int)need);
            }
        }
        // ---- phase C: lists longer than kKeep are produced again by the whole warp, one alignment at a time, one
        //      lane per candidate, straight into match_buf (a lane doing it alone would chain every load) ----
        unsigned big = __ballot_sync(0xffffffffu, live && n > kKeep && fits);
        while (big) {
            const int j = __ffs(big) - 1;
            big &= big - 1;
            // the alignment's data again from (cached) global memory: keeping it in registers until here would spill
            const u64 o_pos = __shfl_sync(0xffffffffu, pos, j);
            const i64 o_a = __shfl_sync(0xffffffffu, a, j);
            const i64 o_w0 = r.aln_walk_off[o_a];
            const int o_L = (int)(r.aln_walk_off[o_a + 1] - o_w0);
            const int o_first = r.walk_node[o_w0], o_last = r.walk_node[o_w0 + o_L - 1];
            const int o_second = o_L > 1 ? r.walk_node[o_w0 + 1] : 0, o_penult = o_L > 1 ? r.walk_node[o_w0 + o_L - 2] : 0;
            const int o_fo = g.occ_off[o_first], o_fcnt = g.occ_off[o_first + 1] - o_fo;
            const int o_ro = g.occ_off[o_last], o_rcnt = o_L > 1 ? g.occ_off[o_last + 1] - o_ro : 0;
            int written = 0;
            for (int c0 = 0; c0 < o_fcnt + o_rcnt; c0 += 32) {          // forward candidates first, then reverse
                const int c = c0 + lane;
                const bool in = c < o_fcnt + o_rcnt;
                const bool rev = c >= o_fcnt;
                bool ok = false;
                int slot = 0;
                if (in) {
                    const int2 e = occ[rev ? o_ro + (c - o_fcnt) : o_fo + c];
                    slot = e.x;
                    ok = o_L == 1 || e.y == (rev ? o_penult : o_second);
                    if (ok && o_L > 2) {
                        const i64 wf = rev ? o_w0 + o_L - 1 : o_w0;
                        const int step = rev ? -1 : 1;
                        int pn[kU], wn[kU];
#pragma unroll
                        for (int u = 0; u < kU; ++u) {
                            pn[u] = wn[u] = 0;
                            if (u + 2 < o_L) {
                                pn[u] = g.slot_node[min(slot + u + 2, last_slot)];
                                wn[u] = r.walk_node[wf + (i64)step * (u + 2)];
                            }
                        }
#pragma unroll
                        for (int u = 0; u < kU; ++u) ok = ok && pn[u] == wn[u];
                        for (int i = kU + 2; i < o_L && ok; ++i)
                            ok = g.slot_node[min(slot + i, last_slot)] == r.walk_node[wf + (i64)step * i];
                    }
                }
                const unsigned ball = __ballot_sync(0xffffffffu, ok);
                if (ok)
                    match_buf[o_pos + written + __popc(ball & below)] =
                        make_uint2((uint32_t)slot | (rev ? SFB_MATCH_REV : 0u), (uint32_t)path_of_slot(g, slot));
                written += __popc(ball);
            }
        }
        __syncwarp();
    }
    const int c0 = warp_sum(st.cand), c1 = warp_sum(st.cmp), c2 = warp_sum(tuples);
    if (lane == 0) {
        SFB_COUNT(sc, SFB_C_ST_CAND, c0);
        SFB_COUNT(sc, SFB_C_ST_COMPARES, c1);
        SFB_COUNT(sc, SFB_C_ST_TUPLES, c2);
    }
    __syncthreads();
    block_counters_flush(sc, counters);
}

static OccCache g_match_cache;
int launch_match(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, cudaStream_t st) {
    if (r->n_alns - w->aln_begin <= 0) return SFB_OK;
    const unsigned grid = persistent_grid(k_match, 256, blocks_for(r->n_alns - w->aln_begin, 256), g_match_cache);   // a warp per 32 alignments
    k_match<<<grid, 256, 0, st>>>(*g, *r, *w, acc->counters, nullptr, nullptr);
    return check_launch("k_match");
}
// the alignments of a device-side list (its length is only known on the device: the grid is the persistent one)
int launch_match_list(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, const int32_t* list,
                      const unsigned long long* n_list, cudaStream_t st) {
    if (r->n_alns - w->aln_begin <= 0) return SFB_OK;
    const unsigned grid = persistent_grid(k_match, 256, blocks_for(r->n_alns - w->aln_begin, 256), g_match_cache);
    k_match<<<grid, 256, 0, st>>>(*g, *r, *w, acc->counters, list, n_list);
    return check_launch("k_match (list)");
}

}  // namespace sfb
