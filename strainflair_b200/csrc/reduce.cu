// Reductions after the two passes:
//   k_path_reduce     per colored path: print_gene_table's statistics (json2csv.py:440-469)
//   k_cluster_counts  per cluster: strain copy counts (print_unassigned_info, json2csv.py:544)
//   k_strain_*        per strain: compute_strains_abundance.py:52-69
// All sums run in a fixed order (strided partials + shuffle/shared-memory trees): results are
// run-to-run identical for given accumulators.
#include "common.cuh"

namespace sfb {

__global__ void __launch_bounds__(256) k_path_reduce(sfb_graph g, sfb_accum acc, i64 p_begin, i64 p_end, double* table) {
    const i64 p = p_begin + ((i64)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
    if (p >= p_end) return;
    const int lo = g.path_off[p], hi = g.path_off[p + 1] - 1;      // the last slot is the sentinel
    double su = 0.0, sum = 0.0;
    int cu = 0, cum = 0, cc = 0;
    for (int s = lo + lane_id(); s < hi; s += kWarp) {
        const double u = acc.uniq_abund[s], m = acc.mult_abund[s];
        const double um = u + m;
        su += u;
        sum += um;
        cu += u > 0.0;
        cum += um > 0.0;
        cc += (u > 0.0) || (m > 0.0);
    }
    su = warp_sum(su);
    sum = warp_sum(sum);
    cu = warp_sum(cu);
    cum = warp_sum(cum);
    cc = warp_sum(cc);
    if (lane_id() == 0) {
        const double n = (double)(hi - lo);
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        double* t = table + (p - p_begin) * SFB_TABLE_COLS;
        t[0] = acc.uniq_norm[p];                                   // nb_uniq_mapped_normalized   (:444)
        t[1] = (double)acc.uniq_reads[p] + acc.mult_reads[p];      // nb_multimapped              (:447)
        t[2] = acc.uniq_norm[p] + acc.mult_norm[p];                // nb_multimapped_normalized   (:450)
        t[3] = su / n;                                             // mean_abund_uniq             (:453)
        t[4] = cu ? su / (double)cu : nan;                         // mean_abund_uniq_nz          (:456)
        t[5] = sum / n;                                            // mean_abund_multiple         (:459)
        t[6] = cum ? sum / (double)cum : nan;                      // mean_abund_multiple_nz      (:462)
        t[7] = (double)cc / n;                                     // ratio_covered_nodes         (:465-469)
    }
}

__global__ void k_cluster_counts(i64 n_clusters, i64 n_strains, const int32_t* clu_off, const int32_t* clu_paths,
                                 const int32_t* pstr_off, const int32_t* pstr_id, const int32_t* pstr_cnt,
                                 long long* out) {
    const i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_clusters) return;
    long long* row = out + c * n_strains;
    for (int j = clu_off[c]; j < clu_off[c + 1]; ++j) {
        const int p = clu_paths[j];
        for (int e = pstr_off[p]; e < pstr_off[p + 1]; ++e) row[pstr_id[e]] += pstr_cnt[e];
    }
}

// ---- strain level ------------------------------------------------------------------------------
// ws_counts: [0,S) sum of counts, [S,2S) sum of counts on covered genes, [2S,3S) number of specific genes,
// [3S,4S) segment offsets.  ws_vals: [0,P) count of the specific strain per row, [P,2P) / [2P,3P) compacted values.
__global__ void k_strain_rows(i64 n_rows, i64 S, const int32_t* row_off, const int32_t* row_strain,
                              const double* row_count, const double* ratio_cov, int32_t* ws_rows, double* ws_vals,
                              double* ws_counts) {
    const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_rows) return;
    int npos = 0, s = -1;
    double c = 0.0;
    for (int e = row_off[p]; e < row_off[p + 1]; ++e)
        if (row_count[e] > 0.0) { ++npos; s = row_strain[e]; c = row_count[e]; }
    if (npos != 1) { ws_rows[p] = -1; return; }      // keep "specific" genes only (:52)
    ws_rows[p] = s;
    ws_vals[p] = c;
    atomicAdd(&ws_counts[s], c);                     // sums of integers: exact, order-free
    if (ratio_cov[p] > 0.0) atomicAdd(&ws_counts[S + s], c);
    atomicAdd(&ws_counts[2 * S + s], 1.0);
}
__global__ void k_strain_offsets(i64 S, double* ws_counts) {
    double run = 0.0;
    for (i64 s = 0; s < S; ++s) { ws_counts[3 * S + s] = run; run += ws_counts[2 * S + s]; }
}

__device__ __forceinline__ u64 ord_key(double x) {
    const u64 b = (u64)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_val(u64 k) {
    const u64 b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
// k-th smallest (0-based) of v[0..n) by MSB-first radix selection; the whole block participates
__device__ double block_select(const double* v, int n, int k, unsigned* s_hist, u64* s_state) {
    u64 prefix = 0, mask = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const u64 key = ord_key(v[i]);
            if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int digit = 0, left = k;
            for (; digit < 255; ++digit) {
                if (left < (int)s_hist[digit]) break;
                left -= (int)s_hist[digit];
            }
            s_state[0] = (u64)digit;
            s_state[1] = (u64)left;
        }
        __syncthreads();
        prefix |= s_state[0] << shift;
        mask |= 0xffull << shift;
        k = (int)s_state[1];
        __syncthreads();
    }
    return ord_val(prefix);
}
__device__ double block_sum(const double* v, int n, double* s_red) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += v[i];
    s_red[threadIdx.x] = acc;
    __syncthreads();
    for (int d = blockDim.x >> 1; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) s_red[threadIdx.x] += s_red[threadIdx.x + d];
        __syncthreads();
    }
    const double out = s_red[0];
    __syncthreads();
    return out;
}

constexpr int kStrainThreads = 256;
__global__ void __launch_bounds__(kStrainThreads) k_strain_stats(i64 n_rows, i64 S, const double* mam,
                                                                 const double* mam_nz, const int32_t* ws_rows,
                                                                 double* ws_vals, const double* ws_counts, double* out) {
    __shared__ int s_scan[kStrainThreads];
    __shared__ unsigned s_hist[256];
    __shared__ u64 s_state[2];
    __shared__ double s_red[kStrainThreads];
    __shared__ int s_base;
    const int s = blockIdx.x;
    const i64 seg = (i64)ws_counts[3 * S + s];
    double* v1 = ws_vals + n_rows + seg;
    double* v2 = ws_vals + 2 * n_rows + seg;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    // ordered compaction of the rows specific to strain s (row order => deterministic sums)
    for (i64 p0 = 0; p0 < n_rows; p0 += kStrainThreads) {
        const i64 p = p0 + threadIdx.x;
        const int flag = (p < n_rows && ws_rows[p] == s) ? 1 : 0;
        s_scan[threadIdx.x] = flag;
        __syncthreads();
        for (int d = 1; d < kStrainThreads; d <<= 1) {
            const int t = (int)threadIdx.x >= d ? s_scan[threadIdx.x - d] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += t;
            __syncthreads();
        }
        if (flag) {
            const int pos = s_base + s_scan[threadIdx.x] - 1;
            v1[pos] = mam[p] / ws_vals[p];       // mean_abund_multiple / count (:62)
            v2[pos] = mam_nz[p] / ws_vals[p];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += s_scan[kStrainThreads - 1];
        __syncthreads();
    }
    const int n = s_base;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    const double tot = ws_counts[s], det = ws_counts[S + s];
    double r[5];
    r[0] = tot == 0.0 ? nan : det / tot;                          // detected_genes (:60)
    if (n == 0) {
        r[1] = r[2] = r[3] = r[4] = nan;
    } else {
        r[1] = block_sum(v1, n, s_red) / (double)n;
        r[2] = block_sum(v2, n, s_red) / (double)n;
        const int k_lo = (n - 1) / 2, k_hi = n / 2;
        const double a1 = block_select(v1, n, k_lo, s_hist, s_state);
        const double b1 = k_hi == k_lo ? a1 : block_select(v1, n, k_hi, s_hist, s_state);
        const double a2 = block_select(v2, n, k_lo, s_hist, s_state);
        const double b2 = k_hi == k_lo ? a2 : block_select(v2, n, k_hi, s_hist, s_state);
        r[3] = 0.5 * (a1 + b1);
        r[4] = 0.5 * (a2 + b2);
    }
    if (threadIdx.x == 0)
        for (int k = 0; k < 5; ++k) out[(i64)s * 5 + k] = r[k];
}
// threshold on detected_genes and normalisation to 100 (:62-69)
__global__ void k_strain_finalize(i64 S, double thr, double* out) {
    const int k = 1 + threadIdx.x;     // one thread per abundance column
    if (k > 4) return;
    double total = 0.0;
    for (i64 s = 0; s < S; ++s) {
        const double det = out[s * 5];
        const double v = det > thr ? out[s * 5 + k] : 0.0;
        out[s * 5 + k] = v;
        if (v == v) total += v;        // pandas sum skips NaN
    }
    if (total != 0.0) {
        const double div = total / 100.0;
        for (i64 s = 0; s < S; ++s) out[s * 5 + k] /= div;
    }
}

int launch_path_reduce(const sfb_graph* g, const sfb_accum* acc, i64 p_begin, i64 p_end, double* table, cudaStream_t st) {
    if (p_begin == p_end) return SFB_OK;
    k_path_reduce<<<blocks_for((p_end - p_begin) * kWarp, 256), 256, 0, st>>>(*g, *acc, p_begin, p_end, table);
    return check_launch("k_path_reduce");
}
int launch_cluster_counts(i64 n_clusters, i64 n_strains, const int32_t* clu_off, const int32_t* clu_paths,
                          const int32_t* pstr_off, const int32_t* pstr_id, const int32_t* pstr_cnt, long long* out,
                          cudaStream_t st) {
    k_cluster_counts<<<blocks_for(n_clusters, 128), 128, 0, st>>>(n_clusters, n_strains, clu_off, clu_paths, pstr_off,
                                                                  pstr_id, pstr_cnt, out);
    return check_launch("k_cluster_counts");
}
int launch_strain_aggregate(i64 n_rows, i64 n_strains, const int32_t* row_off, const int32_t* row_strain,
                            const double* row_count, const double* mam, const double* mam_nz, const double* ratio_cov,
                            double thr, int32_t* ws_rows, double* ws_vals, double* ws_counts, double* out,
                            cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(ws_counts, 0, sizeof(double) * 4 * n_strains, st);
    if (e != cudaSuccess) return fail(SFB_E_CUDA, "memset: %s", cudaGetErrorString(e));
    if (n_rows) {
        k_strain_rows<<<blocks_for(n_rows, 256), 256, 0, st>>>(n_rows, n_strains, row_off, row_strain, row_count,
                                                                ratio_cov, ws_rows, ws_vals, ws_counts);
        SFB_TRY(check_launch("k_strain_rows"));
    }
    k_strain_offsets<<<1, 1, 0, st>>>(n_strains, ws_counts);
    SFB_TRY(check_launch("k_strain_offsets"));
    k_strain_stats<<<(unsigned)n_strains, kStrainThreads, 0, st>>>(n_rows, n_strains, mam, mam_nz, ws_rows, ws_vals,
                                                                    ws_counts, out);
    SFB_TRY(check_launch("k_strain_stats"));
    k_strain_finalize<<<1, 32, 0, st>>>(n_strains, thr, out);
    return check_launch("k_strain_finalize");
}

}  // namespace sfb
