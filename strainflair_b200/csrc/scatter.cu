// Node-level additions:
//   k_pass1_scatter  uniq_abund[slot+i] += cov[i]            (fill_abund_dist, json2csv.py:329-330)
//   k_pass2_scatter  mult_abund[slot+i] += cov[i] * ratio    (json2csv.py:1157-1158, :1209-1210, :1267-1268)
//
// One thread per work item (assigned alignment / pass-2 work item), serial over the item's 5-9 records: the packed
// (aligned, node length) words of a walk are 1-2 sectors, the fp64 REDs of a walk land in 1-2 sectors of the slot
// array, and nothing but the reads' own arrays is touched.  Measured against a tiled record-parallel version
// (shared-memory prefix sums + binary search per record): 1.7x faster on pass 1, see profiles/README.md.
#include "common.cuh"

namespace sfb {

__global__ void __launch_bounds__(256) k_pass1_scatter(sfb_reads r, sfb_work w, double* uniq_abund, i64 det_stride,
                                                       long long* counters) {
    int done = 0;
    for (i64 a = w.aln_begin + (i64)blockIdx.x * blockDim.x + threadIdx.x; a < r.n_alns; a += (i64)gridDim.x * blockDim.x) {
        const int32_t code = w.aln_assign[a];
        if (code < 0) continue;
        const i64 w0 = r.aln_walk_off[a];
        const int L = (int)(r.aln_walk_off[a + 1] - w0);
        double* dst = uniq_abund + ((uint32_t)code & SFB_SLOT_MASK);
        for (int i = 0; i < L; ++i) acc_add(dst + i, det_stride, record_cov(r, w0 + i));
        done += L;
    }
    done = warp_sum(done);
    if (lane_id() == 0 && done) atomicAdd((u64*)&counters[SFB_C_ST_P1_RECORDS], (u64)done);
}
// A work item of the path-set kernels names its path (bit 31 of the code; group_sets.cuh): the slot is where that path
// has the walk's first node -- its last node for a match of the reversed walk -- i.e. occurrence number
// popcount(node_mask & (bit - 1)) of the node (sfb200.h, SFB_MATCH_SETS).
__global__ void __launch_bounds__(256) k_pass2_scatter(sfb_graph g, sfb_reads r, sfb_work w, double* mult_abund, i64 det_stride,
                                                       long long* counters) {
    const App* apps = reinterpret_cast<const App*>(w.app_buf);
    // more work items claimed than app_buf holds: the resolve kernels counted MATCH_OVERFLOW and left claims unwritten;
    // nothing is added (the caller regrows the buffer and reruns the pass)
    const i64 n_app = (i64)w.cursor[2] > (i64)w.app_cap ? 0 : (i64)w.cursor[2];
    int done = 0;
    for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < n_app; q += (i64)gridDim.x * blockDim.x) {
        const App app = apps[q];
        const i64 w0 = r.aln_walk_off[app.aln];
        const int L = (int)(r.aln_walk_off[app.aln + 1] - w0);
        uint32_t slot = app.code & SFB_SLOT_MASK;
        if (app.code & 0x80000000u) {
            const int node = r.walk_node[(app.code & SFB_MATCH_REV) ? w0 + L - 1 : w0];
            const u64 bit = 1ull << ((int)slot - (int)reinterpret_cast<const uint2*>(w.aln_match)[app.aln].y);
            slot = (uint32_t)reinterpret_cast<const int2*>(g.occ_pair)[g.occ_off[node] + __popcll(g.node_mask[node] & (bit - 1))].x;
        }
        double* dst = mult_abund + slot;
        for (int i = 0; i < L; ++i) acc_add(dst + i, det_stride, record_cov(r, w0 + i) * app.weight);
        done += L;
    }
    done = warp_sum(done);
    if (lane_id() == 0 && done) atomicAdd((u64*)&counters[SFB_C_ST_P2_RECORDS], (u64)done);
}

__global__ void k_collapse(double* x, i64 n, i64 stride) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        x[i] = (double)reinterpret_cast<const u64*>(x)[i] * 0x1p-32 + (double)reinterpret_cast<const u64*>(x)[i + stride] * 0x1p-64;
}
// Per-node sums of every path without the sentinel slots, path after path: what Path.unique_mapped_abundances /
// multiple_mapped_abundances hold in the reference (json2csv.py:107-108).  One thread per slot; the slot's path comes
// from the 64-slot block table, its plain index is the slot minus the sentinels before it (one per earlier path).
__global__ void k_drop_sentinels(sfb_graph g, const double* __restrict__ src, double* __restrict__ dst) {
    for (i64 s = (i64)blockIdx.x * blockDim.x + threadIdx.x; s < g.n_slots; s += (i64)gridDim.x * blockDim.x) {
        const int p = path_of_slot(g, (int)s);
        if (s != (i64)g.path_off[p + 1] - 1) dst[s - p] = src[s];
    }
}
int launch_drop_sentinels(const sfb_graph* g, const double* src, double* dst, cudaStream_t st) {
    if (g->n_slots == 0) return SFB_OK;
    k_drop_sentinels<<<kSMs * 8, 256, 0, st>>>(*g, src, dst);
    return check_launch("k_drop_sentinels");
}
int launch_collapse(double* x, i64 n, i64 stride, cudaStream_t st) {
    if (n == 0) return SFB_OK;
    k_collapse<<<kSMs * 8, 256, 0, st>>>(x, n, stride);
    return check_launch("k_collapse");
}

int launch_pass1_scatter(const sfb_graph* g, const sfb_reads* r, const sfb_work* w, sfb_accum* acc, cudaStream_t st) {
    if (r->n_alns - w->aln_begin <= 0) return SFB_OK;
    static OccCache cache;
    const unsigned grid = persistent_grid(k_pass1_scatter, 256, blocks_for(r->n_alns - w->aln_begin, 256), cache);
    k_pass1_scatter<<<grid, 256, 0, st>>>(*r, *w, acc->uniq_abund, acc->det_stride, acc->counters);
    return check_launch("k_pass1_scatter");
}
int launch_pass2_scatter(const sfb_graph* g, const sfb_reads* r, const sfb_work* w, sfb_accum* acc, cudaStream_t st) {
    if (r->n_alns == 0 || w->app_cap == 0) return SFB_OK;
    static OccCache cache;
    const unsigned grid = persistent_grid(k_pass2_scatter, 256, blocks_for(w->app_cap, 256), cache);
    k_pass2_scatter<<<grid, 256, 0, st>>>(*g, *r, *w, acc->mult_abund, acc->det_stride, acc->counters);
    return check_launch("k_pass2_scatter");
}

}  // namespace sfb
