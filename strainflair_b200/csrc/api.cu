// C ABI of libsfb200.so (include/sfb200.h): argument checks and kernel launches.  No allocation, no
// synchronisation, no global mutable state (the last-error string is thread-local).
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace sfb {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SFB_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return SFB_OK;
}

int launch_match(const sfb_graph*, const sfb_reads*, sfb_work*, sfb_accum*, cudaStream_t);
int launch_match_steps(const sfb_graph*, const sfb_reads*, sfb_work*, sfb_accum*, cudaStream_t);
int launch_classify(const Dev&, cudaStream_t);
int launch_pass2_resolve(const Dev&, const long long*, cudaStream_t);
int launch_pass1_scatter(const sfb_graph*, const sfb_reads*, const sfb_work*, sfb_accum*, cudaStream_t);
int launch_pass2_scatter(const sfb_graph*, const sfb_reads*, const sfb_work*, sfb_accum*, cudaStream_t);
int launch_tile_pass(int, const Dev&, const sfb_tiles*, const long long*, cudaStream_t);
int launch_tile_scatter1(const Dev&, const sfb_tiles*, cudaStream_t);
int launch_tile_share(const Dev&, const sfb_tiles*, const long long*, cudaStream_t);
i64 tile_smem_bytes(const sfb_tiles*);
int launch_expand(const sfb_wire*, const int32_t*, const sfb_reads*, i64*, cudaStream_t);
int launch_collapse(double*, i64, i64, cudaStream_t);
int launch_drop_sentinels(const sfb_graph*, const double*, double*, cudaStream_t);
int launch_path_reduce(const sfb_graph*, const sfb_accum*, i64, i64, double*, cudaStream_t);
int launch_cluster_counts(i64, i64, const int32_t*, const int32_t*, const int32_t*, const int32_t*, const int32_t*,
                          long long*, cudaStream_t);
int launch_strain_aggregate(i64, i64, const int32_t*, const int32_t*, const double*, const double*, const double*,
                            const double*, double, int32_t*, double*, double*, double*, cudaStream_t);

static int check_graph(const sfb_graph* g) {
    SFB_REQUIRE(g, "graph is null");
    SFB_REQUIRE(g->n_nodes >= 0 && g->n_paths >= 0 && g->n_slots >= 0 && g->n_strains >= 0, "negative size");
    SFB_REQUIRE(g->n_slots <= SFB_MAX_SLOTS, "n_slots exceeds 2^30-1");
    SFB_REQUIRE(g->mask_words >= 1 && g->mask_words * 64 >= g->n_strains, "mask_words too small");
    SFB_REQUIRE(g->slot_node && g->path_off && g->occ_off && g->occ_pair && g->blk_path &&
                    g->path_seq_len && g->path_nstrain && g->path_strain0 && g->path_smask && g->path_cluster,
                "null graph array");
    return SFB_OK;
}
static int check_reads(const sfb_reads* r) {
    SFB_REQUIRE(r, "reads is null");
    SFB_REQUIRE(r->n_groups >= 0 && r->n_alns >= 0 && r->n_records >= 0 && r->n_exc >= 0, "negative size");
    SFB_REQUIRE(r->n_groups < 0x7fffffffLL && r->n_alns < 0x7fffffffLL, "more than 2^31-1 groups/alignments in one shard");
    SFB_REQUIRE(r->grp_aln_off && r->aln_walk_off, "null reads array");
    SFB_REQUIRE(r->n_groups == 0 || (r->grp_nmates && r->mate_seq_len), "null reads array");
    SFB_REQUIRE(r->n_alns == 0 || (r->aln_bins && r->walk_node && r->walk_alen), "null reads array");
    SFB_REQUIRE(r->n_exc == 0 || r->exc_cov, "exc_cov is null");
    return SFB_OK;
}
static int check_work(const sfb_work* w) {
    SFB_REQUIRE(w, "work is null");
    SFB_REQUIRE(w->aln_match && w->cursor && w->grp_code && w->grp_code2 && w->aln_assign && w->deferred && w->slow,
                "null work array");
    SFB_REQUIRE(w->match_cap >= 0 && w->match_cap < 0x7fffffffLL && (w->match_cap == 0 || w->match_buf),
                "bad match buffer");
    SFB_REQUIRE(w->app_cap >= 0 && (w->app_cap == 0 || w->app_buf), "bad application buffer");
    return SFB_OK;
}
static int check_accum(const sfb_accum* a) {
    SFB_REQUIRE(a, "accum is null");
    SFB_REQUIRE(a->uniq_reads && a->uniq_norm && a->mult_reads && a->mult_norm && a->mult_hits && a->uniq_abund &&
                    a->mult_abund && a->hist && a->counters,
                "null accumulator");
    SFB_REQUIRE(a->det_stride >= 0, "negative det_stride");
    return SFB_OK;
}
static int check_all(const sfb_graph* g, const sfb_reads* r, const sfb_work* w, const sfb_accum* a) {
    SFB_TRY(check_graph(g));
    SFB_TRY(check_reads(r));
    SFB_TRY(check_work(w));
    SFB_REQUIRE(w->grp_begin >= 0 && w->grp_begin <= r->n_groups && w->aln_begin >= 0 && w->aln_begin <= r->n_alns &&
                    w->p2_begin >= 0 && w->p2_begin <= r->n_groups,
                "grp_begin / aln_begin / p2_begin outside the shard");
    return check_accum(a);
}

}  // namespace sfb

using namespace sfb;

extern "C" {

int sfb_abi_version(void) { return SFB_ABI_VERSION; }
const char* sfb_version(void) { return "sfb200 0.2 (sm_100a)"; }
const char* sfb_last_error(void) { return g_err; }

int64_t sfb_workspace_bytes(int what, const sfb_dims* d) {
    if (!d || d->n_groups < 0 || d->n_alns < 0 || d->n_records < 0 || d->n_paths < 0 || d->n_slots < 0 || d->n_strains < 0)
        return fail(SFB_E_ARG, "sfb_workspace_bytes: null or negative dimensions");
    const int64_t G = d->n_groups, A = d->n_alns, R = d->n_records, P = d->n_paths, T = d->n_slots, S = d->n_strains;
    auto al = [](int64_t n) { return (n + 255) & ~(int64_t)255; };
    switch (what) {
    case SFB_WS_EXPANDED:
        return al(8 * (2 * G + 1)) + al(8 * (A + 1)) + al(8 * ((2 * G > A ? 2 * G : A) / 4096 + 2)) + al(8 * G) + al(5 * A) +
               al(4 * R) + al(4 * R);
    case SFB_WS_MATCH_PAIRS: return 2 * A + 1024;
    case SFB_WS_APP_ITEMS: return A + (2 * A + 1024);
    case SFB_WS_ACCUM_F64: return (d->deterministic ? 2 : 1) * (3 * P + 2 * T) + 1;
    case SFB_WS_ACCUM_I64: return P + (int64_t)SFB_N_HIST * S * SFB_N_BINS + SFB_N_COUNTERS + (P + 1) / 2;
    default: return fail(SFB_E_ARG, "sfb_workspace_bytes: unknown size %d", what);
    }
}

int sfb_expand_reads(const sfb_wire* w, const int32_t* node_len, const sfb_reads* dst, int64_t* scratch, void* stream) {
    SFB_REQUIRE(w && dst, "wire or destination is null");
    SFB_REQUIRE(w->n_groups >= 0 && w->n_alns >= 0 && w->n_records >= 0 && w->n_exc >= 0 && w->n_wide >= 0 &&
                    w->n_seq_exc >= 0 && w->n_bins_exc >= 0, "negative size");
    SFB_REQUIRE(dst->grp_aln_off && dst->aln_walk_off && scratch, "null output");
    SFB_REQUIRE(w->n_groups == 0 || (w->grp_nmates && w->mate_naln && dst->mate_seq_len), "null array");
    SFB_REQUIRE(w->n_seq_exc == 0 || (w->seq_exc_idx && w->seq_exc_val), "null array");
    SFB_REQUIRE(w->n_alns == 0 || (w->aln_len && w->aln_code && w->bins_dict && w->aln_node0 && w->node0_anchor && w->aln_al_first &&
                                   w->aln_al_last && dst->aln_bins), "null array");
    SFB_REQUIRE(w->n_bins_exc == 0 || (w->bins_exc_idx && w->bins_exc_val), "null array");
    SFB_REQUIRE(w->n_records == 0 || (w->walk_step && dst->walk_node && dst->walk_alen && node_len), "null array");
    SFB_REQUIRE(w->n_exc == 0 || (w->exc_idx && w->exc_cov), "null array");
    SFB_REQUIRE(w->n_wide == 0 || (w->wide_idx && w->wide_node), "null array");
    return launch_expand(w, node_len, dst, (i64*)scratch, (cudaStream_t)stream);
}

int sfb_match(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, void* stream) {
    SFB_TRY(check_all(g, r, w, acc));
    return launch_match(g, r, w, acc, (cudaStream_t)stream);
}

int sfb_match_steps(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, void* stream) {
    SFB_TRY(check_all(g, r, w, acc));
    SFB_REQUIRE(g->n_nodes == 0 || (g->node_step && g->node_mask && g->step_off), "the graph has no step index");
    SFB_REQUIRE(((uintptr_t)g->node_step & 31) == 0, "node_step must be 32-byte aligned");
    SFB_REQUIRE(w->aln_sets && ((uintptr_t)w->aln_sets & 15) == 0, "aln_sets is null or not 16-byte aligned");
    return launch_match_steps(g, r, w, acc, (cudaStream_t)stream);
}

int sfb_classify(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, void* stream) {
    SFB_TRY(check_all(g, r, w, acc));
    Dev d{*g, *r, *w, *acc};
    return launch_classify(d, (cudaStream_t)stream);
}

int sfb_tile_pass1(const sfb_graph* g, const sfb_tiles* t, const sfb_reads* r, sfb_work* w, sfb_accum* acc, void* stream) {
    SFB_TRY(check_all(g, r, w, acc));
    Dev d{*g, *r, *w, *acc};
    return launch_tile_pass(1, d, t, nullptr, (cudaStream_t)stream);
}

int sfb_tile_pass2(const sfb_graph* g, const sfb_tiles* t, const sfb_reads* r, sfb_work* w, sfb_accum* acc,
                   const long long* uniq_reads_global, void* stream) {
    SFB_TRY(check_all(g, r, w, acc));
    SFB_REQUIRE(uniq_reads_global, "uniq_reads_global is null");
    Dev d{*g, *r, *w, *acc};
    return launch_tile_pass(2, d, t, uniq_reads_global, (cudaStream_t)stream);
}

int sfb_tile_scatter1(const sfb_graph* g, const sfb_tiles* t, const sfb_reads* r, sfb_work* w, sfb_accum* acc, void* stream) {
    SFB_TRY(check_all(g, r, w, acc));
    Dev d{*g, *r, *w, *acc};
    return launch_tile_scatter1(d, t, (cudaStream_t)stream);
}

int sfb_tile_share(const sfb_graph* g, const sfb_tiles* t, const sfb_reads* r, sfb_work* w, sfb_accum* acc,
                   const long long* uniq_reads_global, void* stream) {
    SFB_TRY(check_all(g, r, w, acc));
    SFB_REQUIRE(uniq_reads_global, "uniq_reads_global is null");
    SFB_REQUIRE(g->node_mask && w->aln_sets, "sfb_tile_share works on the path sets of sfb_match_steps");
    Dev d{*g, *r, *w, *acc};
    return launch_tile_share(d, t, uniq_reads_global, (cudaStream_t)stream);
}

int64_t sfb_tile_smem_bytes(const sfb_tiles* t) { return tile_smem_bytes(t); }

int sfb_pass1_scatter(const sfb_graph* g, const sfb_reads* r, const sfb_work* w, sfb_accum* acc, void* stream) {
    SFB_TRY(check_all(g, r, w, acc));
    return launch_pass1_scatter(g, r, w, acc, (cudaStream_t)stream);
}

int sfb_pass2(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, const long long* uniq_reads_global,
              void* stream) {
    SFB_TRY(check_all(g, r, w, acc));
    SFB_REQUIRE(uniq_reads_global, "uniq_reads_global is null");
    Dev d{*g, *r, *w, *acc};
    SFB_TRY(launch_pass2_resolve(d, uniq_reads_global, (cudaStream_t)stream));
    return launch_pass2_scatter(g, r, w, acc, (cudaStream_t)stream);
}

int sfb_pass2_resolve(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc,
                      const long long* uniq_reads_global, void* stream) {
    SFB_TRY(check_all(g, r, w, acc));
    SFB_REQUIRE(uniq_reads_global, "uniq_reads_global is null");
    Dev d{*g, *r, *w, *acc};
    return launch_pass2_resolve(d, uniq_reads_global, (cudaStream_t)stream);
}

int sfb_pass2_scatter(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, void* stream) {
    SFB_TRY(check_all(g, r, w, acc));
    return launch_pass2_scatter(g, r, w, acc, (cudaStream_t)stream);
}

int sfb_accum_collapse(double* x, int64_t n, int64_t det_stride, void* stream) {
    SFB_REQUIRE(n >= 0 && det_stride >= n, "bad sizes");
    SFB_REQUIRE(x || n == 0, "null array");
    return launch_collapse(x, n, det_stride, (cudaStream_t)stream);
}

int sfb_drop_sentinels(const sfb_graph* g, const double* slots, double* plain, void* stream) {
    SFB_TRY(check_graph(g));
    SFB_REQUIRE((slots && plain) || g->n_slots == 0, "null array");
    return launch_drop_sentinels(g, slots, plain, (cudaStream_t)stream);
}

int sfb_path_reduce(const sfb_graph* g, const sfb_accum* acc, int64_t p_begin, int64_t p_end, double* table,
                    void* stream) {
    SFB_TRY(check_graph(g));
    SFB_TRY(check_accum(acc));
    SFB_REQUIRE(0 <= p_begin && p_begin <= p_end && p_end <= g->n_paths, "bad path range");
    SFB_REQUIRE(table || p_begin == p_end, "table is null");
    return launch_path_reduce(g, acc, p_begin, p_end, table, (cudaStream_t)stream);
}

int sfb_cluster_strain_counts(int64_t n_clusters, int64_t n_strains, const int32_t* clu_off, const int32_t* clu_paths,
                              const int32_t* pstr_off, const int32_t* pstr_id, const int32_t* pstr_cnt, long long* out,
                              void* stream) {
    SFB_REQUIRE(n_clusters >= 0 && n_strains >= 0, "negative size");
    if (n_clusters == 0 || n_strains == 0) return SFB_OK;
    SFB_REQUIRE(clu_off && clu_paths && pstr_off && pstr_id && pstr_cnt && out, "null array");
    return launch_cluster_counts(n_clusters, n_strains, clu_off, clu_paths, pstr_off, pstr_id, pstr_cnt, out,
                                 (cudaStream_t)stream);
}

int sfb_strain_aggregate(int64_t n_rows, int64_t n_strains, const int32_t* row_off, const int32_t* row_strain,
                         const double* row_count, const double* mam, const double* mam_nz, const double* ratio_cov,
                         double thr, int32_t* ws_rows, double* ws_vals, double* ws_counts, double* out, void* stream) {
    SFB_REQUIRE(n_rows >= 0 && n_strains >= 0, "negative size");
    if (n_strains == 0) return SFB_OK;
    SFB_REQUIRE(out && ws_counts, "null output");
    // row_strain / row_count may be null when no row has a positive count
    SFB_REQUIRE(n_rows == 0 || (row_off && mam && mam_nz && ratio_cov && ws_rows && ws_vals), "null array");
    return launch_strain_aggregate(n_rows, n_strains, row_off, row_strain, row_count, mam, mam_nz, ratio_cov, thr,
                                   ws_rows, ws_vals, ws_counts, out, (cudaStream_t)stream);
}

}  // extern "C"
