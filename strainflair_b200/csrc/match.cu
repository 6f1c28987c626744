// k_match: Pangenome.get_matching_path (json2csv.py:341-363) for the walk and, if longer than one node,
// the reversed walk (callers :855-866, :1024-1026, :1103-1114, :1231-1233).
//
// One thread per retained alignment.  The node -> slot occurrence index replaces Node.traversed_path and
// the linear position scan; a candidate is rejected at its first differing node, and the sentinel slot
// after every path rejects walks that run off the path end.  The index and the path slots are L2-resident
// gathers (tens of MB), the walk arrays stream from HBM once.  Output: 8 bytes per alignment, plus one
// (code, path id) pair per match for alignments with several matches, in space claimed with one atomic
// per warp.
#include "common.cuh"

namespace sfb {

constexpr int kKeep = 8;   // matches kept in registers; longer lists are produced by a second sweep

struct MatchStats {
    int cand, cmp;
};

template <bool kEmit>
__device__ __forceinline__ int match_sweep(const sfb_graph& g, const sfb_reads& r, i64 w0, int L, uint2* keep,
                                           uint2* out, MatchStats& st) {
    int n = 0;
    const int first = r.walk_node[w0];
    for (int o = g.occ_off[first], oe = g.occ_off[first + 1]; o < oe; ++o) {
        const int s = g.occ_slot[o];
        int i = 1;
        for (; i < L; ++i)
            if (g.slot_node[s + i] != r.walk_node[w0 + i]) break;      // sentinel (-1) stops overruns
        st.cand += 1;
        st.cmp += min(i, L - 1);
        if (i >= L) {
            const uint2 t = make_uint2((uint32_t)s, (uint32_t)g.slot_path[s]);
            if (kEmit) out[n] = t; else if (n < kKeep) keep[n] = t;
            ++n;
        }
    }
    if (L > 1) {   // reversed walk: occurrences of walk[L-1], then walk[L-2] ...
        const int last = r.walk_node[w0 + L - 1];
        for (int o = g.occ_off[last], oe = g.occ_off[last + 1]; o < oe; ++o) {
            const int s = g.occ_slot[o];
            int i = 1;
            for (; i < L; ++i)
                if (g.slot_node[s + i] != r.walk_node[w0 + L - 1 - i]) break;
            st.cand += 1;
            st.cmp += min(i, L - 1);
            if (i >= L) {
                const uint2 t = make_uint2((uint32_t)s | SFB_MATCH_REV, (uint32_t)g.slot_path[s]);
                if (kEmit) out[n] = t; else if (n < kKeep) keep[n] = t;
                ++n;
            }
        }
    }
    return n;
}

__global__ void __launch_bounds__(256) k_match(sfb_graph g, sfb_reads r, sfb_work w, long long* counters) {
    __shared__ BlockCounters sc;
    block_counters_zero(sc);
    __syncthreads();
    MatchStats st{0, 0};
    int tuples = 0;
    uint2* aln_match = reinterpret_cast<uint2*>(w.aln_match);
    uint2* match_buf = reinterpret_cast<uint2*>(w.match_buf);
    const i64 n_round = (r.n_alns + 255) & ~(i64)255;
    for (i64 a = (i64)blockIdx.x * blockDim.x + threadIdx.x; a < n_round; a += (i64)gridDim.x * blockDim.x) {
        const bool live = a < r.n_alns;
        uint2 keep[kKeep];
        int n = 0, L = 0;
        i64 w0 = 0;
        if (live) {
            w0 = r.aln_walk_off[a];
            L = (int)(r.aln_walk_off[a + 1] - w0);
            if (L > 0) n = match_sweep<false>(g, r, w0, L, keep, nullptr, st);
            tuples += n;
        }
        const unsigned need = n > 1 ? (unsigned)n : 0u;
        const u64 pos = warp_claim(w.cursor, need);
        if (!live) continue;
        if (n == 0) {
            aln_match[a] = make_uint2(SFB_MATCH_NONE, 0u);
        } else if (n == 1) {
            aln_match[a] = keep[0];
        } else if (pos + need <= (u64)w.match_cap) {
            if (n <= kKeep) {
                for (int k = 0; k < n; ++k) match_buf[pos + k] = keep[k];
            } else {
                MatchStats scratch{0, 0};
                match_sweep<true>(g, r, w0, L, nullptr, match_buf + pos, scratch);
            }
            aln_match[a] = make_uint2(SFB_MATCH_MULTI | (uint32_t)pos, (uint32_t)n);
        } else {
            aln_match[a] = make_uint2(SFB_MATCH_NONE, 0u);
            SFB_COUNT(sc, SFB_C_MATCH_OVERFLOW, (int)need);
        }
    }
    const int c0 = warp_sum(st.cand), c1 = warp_sum(st.cmp), c2 = warp_sum(tuples);
    if (lane_id() == 0) {
        SFB_COUNT(sc, SFB_C_ST_CAND, c0);
        SFB_COUNT(sc, SFB_C_ST_COMPARES, c1);
        SFB_COUNT(sc, SFB_C_ST_TUPLES, c2);
    }
    __syncthreads();
    block_counters_flush(sc, counters);
}

int launch_match(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, cudaStream_t st) {
    if (r->n_alns == 0) return SFB_OK;
    const unsigned grid = (unsigned)min((i64)kSMs * 8, (i64)blocks_for(r->n_alns, 256));
    k_match<<<grid, 256, 0, st>>>(*g, *r, *w, acc->counters);
    return check_launch("k_match");
}

}  // namespace sfb
