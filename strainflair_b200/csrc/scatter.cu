// Node-level additions, record-parallel:
//   k_pass1_scatter  uniq_abund[slot+i] += cov[i]            (fill_abund_dist, json2csv.py:329-330)
//   k_pass2_scatter  mult_abund[slot+i] += cov[i] * ratio    (json2csv.py:1157-1158, :1209-1210, :1267-1268)
//
// A block takes a tile of 256 work items (alignments / pass-2 applications), builds the prefix sum of
// their walk lengths in shared memory and then walks the tile's records with consecutive threads on
// consecutive records: walk_alen is read coalesced, the node lengths and the fp64 REDs of one walk fall
// in one or two 32-byte sectors, and no thread serialises a whole walk.
#include "common.cuh"

namespace sfb {

constexpr int kTile = 256;

// inclusive scan of one int per thread over a 256-thread block; s_warp needs 8 ints
__device__ __forceinline__ int block_scan_256(int v, int* s_warp) {
    const unsigned incl = warp_scan((unsigned)v);
    if (lane_id() == 31) s_warp[threadIdx.x >> 5] = (int)incl;
    __syncthreads();
    int base = 0;
    for (int wi = 0; wi < (int)(threadIdx.x >> 5); ++wi) base += s_warp[wi];
    __syncthreads();
    return base + (int)incl;
}

__device__ __forceinline__ int find_item(const int* s_off, int n, int k) {
    int lo = 0, hi = n;                    // largest j with s_off[j] <= k
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_off[mid] <= k) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kTile) k_pass1_scatter(sfb_graph g, sfb_reads r, sfb_work w, double* uniq_abund,
                                                         long long* counters) {
    __shared__ int s_off[kTile + 1];
    __shared__ int32_t s_code[kTile];
    __shared__ int s_done;
    if (threadIdx.x == 0) s_done = 0;
    __syncthreads();
    int done = 0;
    for (i64 tile = blockIdx.x; tile * kTile < r.n_alns; tile += gridDim.x) {
        const i64 a_lo = tile * kTile;
        const int na = (int)min((i64)kTile, r.n_alns - a_lo);
        const i64 r_lo = r.aln_walk_off[a_lo];
        for (int j = threadIdx.x; j <= na; j += blockDim.x) s_off[j] = (int)(r.aln_walk_off[a_lo + j] - r_lo);
        for (int j = threadIdx.x; j < na; j += blockDim.x) s_code[j] = w.aln_assign[a_lo + j];
        __syncthreads();
        const int nrec = s_off[na];
        for (int k = threadIdx.x; k < nrec; k += blockDim.x) {
            const int j = find_item(s_off, na, k);
            const int32_t code = s_code[j];
            if (code < 0) continue;
            const int i = k - s_off[j];
            const int L = s_off[j + 1] - s_off[j];
            const double cov = record_cov(g, r, r_lo + k, (uint32_t)code, i, L);
            atomicAdd(&uniq_abund[((uint32_t)code & SFB_SLOT_MASK) + i], cov);
            ++done;
        }
        __syncthreads();
    }
    done = warp_sum(done);
    if (lane_id() == 0 && done) atomicAdd(&s_done, done);
    __syncthreads();
    if (threadIdx.x == 0 && s_done) atomicAdd((u64*)&counters[SFB_C_ST_P1_RECORDS], (u64)s_done);
}

__global__ void __launch_bounds__(kTile) k_pass2_scatter(sfb_graph g, sfb_reads r, sfb_work w, double* mult_abund,
                                                         long long* counters) {
    __shared__ int s_off[kTile + 1];
    __shared__ int s_warp[8];
    __shared__ i64 s_w0[kTile];
    __shared__ uint32_t s_code[kTile];
    __shared__ double s_weight[kTile];
    __shared__ int s_done;
    if (threadIdx.x == 0) { s_done = 0; s_off[0] = 0; }
    __syncthreads();
    const App* apps = reinterpret_cast<const App*>(w.app_buf);
    const i64 n_app = min((i64)w.cursor[2], (i64)w.app_cap);
    int done = 0;
    for (i64 tile = blockIdx.x; tile * kTile < n_app; tile += gridDim.x) {
        const i64 q = tile * kTile + threadIdx.x;
        int L = 0;
        if (q < n_app) {
            const App app = apps[q];
            const i64 w0 = r.aln_walk_off[app.aln];
            L = (int)(r.aln_walk_off[app.aln + 1] - w0);
            s_w0[threadIdx.x] = w0;
            s_code[threadIdx.x] = app.code;
            s_weight[threadIdx.x] = app.weight;
        }
        s_off[threadIdx.x + 1] = block_scan_256(L, s_warp);
        __syncthreads();
        const int n_items = (int)min((i64)kTile, n_app - tile * kTile);
        const int nrec = s_off[n_items];
        for (int k = threadIdx.x; k < nrec; k += blockDim.x) {
            const int j = find_item(s_off, n_items, k);
            const int i = k - s_off[j];
            const int Lj = s_off[j + 1] - s_off[j];
            const uint32_t code = s_code[j];
            const double cov = record_cov(g, r, s_w0[j] + i, code, i, Lj);
            atomicAdd(&mult_abund[(code & SFB_SLOT_MASK) + i], cov * s_weight[j]);
        }
        done += nrec > (int)threadIdx.x ? (nrec - (int)threadIdx.x + kTile - 1) / kTile : 0;
        __syncthreads();
    }
    done = warp_sum(done);
    if (lane_id() == 0 && done) atomicAdd(&s_done, done);
    __syncthreads();
    if (threadIdx.x == 0 && s_done) atomicAdd((u64*)&counters[SFB_C_ST_P2_RECORDS], (u64)s_done);
}

int launch_pass1_scatter(const sfb_graph* g, const sfb_reads* r, const sfb_work* w, sfb_accum* acc, cudaStream_t st) {
    if (r->n_alns == 0) return SFB_OK;
    static int cache = 0;
    const unsigned grid = persistent_grid(k_pass1_scatter, kTile, blocks_for(r->n_alns, kTile), &cache);
    k_pass1_scatter<<<grid, kTile, 0, st>>>(*g, *r, *w, acc->uniq_abund, acc->counters);
    return check_launch("k_pass1_scatter");
}
int launch_pass2_scatter(const sfb_graph* g, const sfb_reads* r, const sfb_work* w, sfb_accum* acc, cudaStream_t st) {
    if (r->n_alns == 0 || w->app_cap == 0) return SFB_OK;
    static int cache = 0;
    const unsigned grid = persistent_grid(k_pass2_scatter, kTile, blocks_for(w->app_cap, kTile), &cache);
    k_pass2_scatter<<<grid, kTile, 0, st>>>(*g, *r, *w, acc->mult_abund, acc->counters);
    return check_launch("k_pass2_scatter");
}

}  // namespace sfb
