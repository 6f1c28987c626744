// Effects of the per-group logic on the accumulators in HBM, shared by the group kernels (group.cu) and the kernels
// that keep the node-level sums in shared memory (tilesets.cu).
#pragma once
#include "common.cuh"

namespace sfb {

// ---- shared pieces --------------------------------------------------------------------------
__device__ __forceinline__ void hist_add(const Dev& d, BlockCounters& sc, int pid, i64 a) {
    if (d.g.path_nstrain[pid] != 1) return;                       // json2csv.py:333 / :1246
    const int s = d.g.path_strain0[pid];
    SFB_COUNT(sc, SFB_C_ST_HIST, 1);
    for (int k = 0; k < SFB_N_HIST; ++k) {
        const int b = d.r.aln_bins[a * SFB_N_HIST + k];
        if (b >= SFB_N_BINS) { SFB_COUNT(sc, SFB_C_BAD_BIN, 1); continue; }
        atomicAdd(&d.acc.hist[((size_t)k * d.g.n_strains + s) * SFB_N_BINS + b], 1ull);
    }
}

// fill_abund_dist, per-read part (json2csv.py:326-327, 333-339); the node loop (:329-330) is k_pass1_scatter
__device__ __forceinline__ void assign_unique(const Dev& d, BlockCounters& sc, i64 a, uint32_t code, int pid,
                                           int seq_len) {
    d.w.aln_assign[a] = (int32_t)code;
    atomicAdd((u64*)&d.acc.uniq_reads[pid], 1ull);
    acc_add(&d.acc.uniq_norm[pid], d.acc.det_stride, (double)seq_len / (double)d.g.path_seq_len[pid]);
    hist_add(d, sc, pid, a);
    if (a < d.w.aln_begin) {
        // a read group of the tile range that sfb_tile_pass1 handed over (more match tuples than its buffers hold):
        // k_pass1_scatter does not visit it, the node loop (json2csv.py:329-330) runs here
        const i64 w0 = d.r.aln_walk_off[a];
        const int L = (int)(d.r.aln_walk_off[a + 1] - w0);
        double* dst = d.acc.uniq_abund + (code & SFB_SLOT_MASK);
        for (int i = 0; i < L; ++i) acc_add(dst + i, d.acc.det_stride, record_cov(d.r, w0 + i));
        SFB_COUNT(sc, SFB_C_ST_P1_RECORDS, L);
    }
}

// cluster of the first path through the walk's first node (json2csv.py:875,901,1036); -1 = None, -2 = no path
__device__ __forceinline__ int first_cluster(const Dev& d, i64 a) {
    const int n0 = d.r.walk_node[d.r.aln_walk_off[a]];
    const int o = d.g.occ_off[n0];
    if (o == d.g.occ_off[n0 + 1]) return -2;
    return d.g.path_cluster[path_of_slot(d.g, d.g.occ_pair[2 * (size_t)o])];
}
__device__ __forceinline__ void book(const Dev& d, BlockCounters& sc, const Mate& mt, int cluster) {
    if (cluster < 0) { SFB_COUNT(sc, SFB_C_NO_CLUSTER, 1); return; }
    d.w.aln_assign[mt.a0] = -(cluster + 2);
}

// json2csv.py:1155-1156 (and :1207-1208, :1265-1266) + the work item for the node loop (:1157-1158)
__device__ __forceinline__ void emit_app(const Dev& d, App* out, i64 a, uint32_t code, int pid, int seq_len,
                                      double ratio, int times) {
    acc_add(&d.acc.mult_reads[pid], d.acc.det_stride, ratio * (double)times);
    acc_add(&d.acc.mult_norm[pid], d.acc.det_stride,
            (ratio * (double)seq_len / (double)d.g.path_seq_len[pid]) * (double)times);
    d.acc.mult_hits[pid] = 1u;           // "touched" flag: same value from every writer, no atomic needed
    App app;
    app.aln = (int32_t)a;
    app.code = code;
    app.weight = ratio * (double)times;
    *out = app;
}

}  // namespace sfb
