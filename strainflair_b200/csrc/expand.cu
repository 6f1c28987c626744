// sfb_expand_reads: compact transfer ("wire") layout of a read shard -> the arrays the kernels consume.
//
// The end-to-end path pays the host->device copy (PCIe, ~55 GB/s): CSR offsets travel as 1-byte counts and become int64
// offsets by a device prefix sum; read lengths travel as 16 bits; per record four bits travel: the step from the previous
// node id of the walk (walks follow the graph, ids are locally consecutive); aligned bases only for the two end records of
// a walk (the others cover their node); histogram bins as a dictionary index; one common read length.  Node ids are
// rebuilt by summing the steps along every walk, node lengths are gathered from the graph on the device.
// 3.7 GB -> 0.56 GB per C3 shard (first nodes: 16 bits above their block's smallest one).
#include "common.cuh"

namespace sfb {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 16;                       // per thread
constexpr int kScanTile = kScanBlock * kScanItems;   // 4096 counts per CTA

__device__ __forceinline__ i64 block_excl_scan(i64 v, i64* s_warp, i64& block_total) {
    // exclusive scan of one value per thread over a 256-thread block
    i64 incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const i64 t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane_id() >= d) incl += t;
    }
    if (lane_id() == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    i64 base = 0, total = 0;
    for (int wi = 0; wi < kScanBlock / 32; ++wi) {
        if (wi < (int)(threadIdx.x >> 5)) base += s_warp[wi];
        total += s_warp[wi];
    }
    __syncthreads();
    block_total = total;
    return base + incl - v;
}

// pass 1: sum of each tile of 4096 counts
__global__ void __launch_bounds__(kScanBlock) k_scan_partials(const uint8_t* cnt, i64 n, i64* partial) {
    __shared__ i64 s_warp[kScanBlock / 32];
    const i64 lo = (i64)blockIdx.x * kScanTile + (i64)threadIdx.x * kScanItems;
    i64 s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (lo + i < n) s += cnt[lo + i];
    i64 total;
    block_excl_scan(s, s_warp, total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}
// pass 2: exclusive scan of the tile sums, one CTA
__global__ void __launch_bounds__(kScanBlock) k_scan_tiles(i64* partial, i64 n_tiles) {
    __shared__ i64 s_warp[kScanBlock / 32];
    i64 carry = 0;
    for (i64 base = 0; base < n_tiles; base += kScanBlock) {
        const i64 i = base + threadIdx.x;
        const i64 v = i < n_tiles ? partial[i] : 0;
        i64 total;
        const i64 ex = block_excl_scan(v, s_warp, total);
        if (i < n_tiles) partial[i] = carry + ex;
        carry += total;
    }
}
// pass 3: offsets inside the tile + tile base; off[n] = total
__global__ void __launch_bounds__(kScanBlock) k_scan_write(const uint8_t* cnt, i64 n, const i64* partial, i64* off) {
    __shared__ i64 s_warp[kScanBlock / 32];
    const i64 lo = (i64)blockIdx.x * kScanTile + (i64)threadIdx.x * kScanItems;
    int c[kScanItems];
    i64 s = 0;
    // a thread's 16 counts with one 128-bit load, its 16 offsets with eight 128-bit stores (a warp's 8-byte stores, 128
    // bytes apart, touch 32 sectors for 256 bytes) -- whenever the arrays are 16-byte aligned and the thread's stretch is whole
    const bool wide = lo + kScanItems <= n && (((size_t)cnt | (size_t)off) & 15) == 0;
    if (wide) {
        const uint4 q = *reinterpret_cast<const uint4*>(cnt + lo);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            c[i] = (int)((w[i >> 2] >> (8 * (i & 3))) & 0xffu);
            s += c[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            c[i] = lo + i < n ? cnt[lo + i] : 0;
            s += c[i];
        }
    }
    i64 total;
    i64 run = partial[blockIdx.x] + block_excl_scan(s, s_warp, total);
    if (wide) {
#pragma unroll
        for (int i = 0; i < kScanItems; i += 2) {
            const i64 a = run, b = run + c[i];
            *reinterpret_cast<longlong2*>(off + lo + i) = make_longlong2(a, b);
            run = b + c[i + 1];
        }
        if (lo + kScanItems == n) off[n] = run;
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) {
            if (lo + i < n) off[lo + i] = run;
            run += c[i];
            if (lo + i == n - 1) off[n] = run;
        }
    }
}

__global__ void k_expand_seq_len(sfb_wire w, int32_t* out) {
    const i64 n = 2 * w.n_groups;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        out[i] = w.mate_seq_len ? (int32_t)w.mate_seq_len[i] : ((i & 1) < w.grp_nmates[i >> 1] ? (int32_t)w.seq_len_common : 0);
}
__global__ void k_apply_seq_exceptions(sfb_wire w, int32_t* out) {
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < w.n_seq_exc; k += (i64)gridDim.x * blockDim.x)
        out[w.seq_exc_idx[k]] = w.seq_exc_val[k];
}
// histogram bins of every alignment from the shard's dictionary of 5-tuples
__global__ void __launch_bounds__(256) k_expand_bins(sfb_wire w, uint8_t* out) {
    __shared__ uint8_t s_dict[256 * SFB_N_HIST];
    for (int i = threadIdx.x; i < 255 * SFB_N_HIST; i += blockDim.x) s_dict[i] = w.bins_dict[i];
    __syncthreads();
    for (i64 a = (i64)blockIdx.x * blockDim.x + threadIdx.x; a < w.n_alns; a += (i64)gridDim.x * blockDim.x) {
        const int code = w.aln_code[a];
#pragma unroll
        for (int k = 0; k < SFB_N_HIST; ++k) out[a * SFB_N_HIST + k] = code < 255 ? s_dict[code * SFB_N_HIST + k] : (uint8_t)255;
    }
}
__global__ void k_apply_bins_exceptions(sfb_wire w, uint8_t* out) {
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < w.n_bins_exc; k += (i64)gridDim.x * blockDim.x)
        for (int q = 0; q < SFB_N_HIST; ++q) out[w.bins_exc_idx[k] * SFB_N_HIST + q] = w.bins_exc_val[k * SFB_N_HIST + q];
}
// Node ids and packed coverage words (aligned bases | node length << 16, see record_cov) of every record.
// A block takes 256 consecutive walks = one contiguous stretch of records.  One thread per walk sums the steps along
// its walk (a handful of records) and gathers the node lengths, writing into a shared-memory image of the stretch;
// the block then stores the image with coalesced 4-byte rows (a thread writing its own 6 records straight to global
// memory touches 32 sectors per store instruction).  Stretches that do not fit the image are written directly.
constexpr int kWalkTile = 256;          // walks per block pass
constexpr int kWalkImage = 2560;        // records of the shared-memory image (mean walk: 5-7 records)

__device__ __forceinline__ int32_t wide_node_of(const sfb_wire& w, i64 r) {
    i64 x = 0, y = w.n_wide;            // wide step: binary search of the record in the (short) list
    while (x < y) {
        const i64 mid = (x + y) >> 1;
        if (w.wide_idx[mid] < r) x = mid + 1;
        else y = mid;
    }
    return w.wide_node[x];
}

__global__ void __launch_bounds__(kWalkTile) k_expand_walks(sfb_wire w, const i64* __restrict__ walk_off,
                                                            const int32_t* __restrict__ node_len,
                                                            int32_t* __restrict__ walk_node, int32_t* __restrict__ walk_alen) {
    __shared__ int32_t s_node[kWalkImage];
    __shared__ int32_t s_alen[kWalkImage];
    const i64 n_tiles = (w.n_alns + kWalkTile - 1) / kWalkTile;
    for (i64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const i64 a0 = tile * kWalkTile;
        const i64 a1 = a0 + kWalkTile < (i64)w.n_alns ? a0 + kWalkTile : (i64)w.n_alns;
        const i64 r0 = walk_off[a0], r1 = walk_off[a1];
        const bool staged = r1 - r0 <= kWalkImage;
        const i64 a = a0 + threadIdx.x;
        if (a < a1) {
            const i64 lo = walk_off[a];
            const int L = (int)w.aln_len[a];                     // (<= 255: everything below in 32-bit arithmetic)
            const int32_t anchor = w.node0_anchor[a >> 8];       // (kWalkTile == 256: one anchor per block pass)
            int32_t node = anchor >= 0 ? anchor + (int32_t)w.aln_node0[a] : w.node0_wide[256 * (i64)(-1 - anchor) + (a & 255)];
            const int32_t al_first = w.aln_al_first[a], al_last = w.aln_al_last[a];
            const int base = staged ? (int)(lo - r0) : 0;        // the walk's place in the shared-memory image
            const uint8_t* sp = w.walk_step + (lo >> 1);         // its nibbles: number odd .. odd + L - 1 from here
            const int odd = (int)(lo & 1);
            // eight records at a time: their nibbles (five bytes at most), then their node lengths, are loaded together (a
            // record-by-record loop would chain two dependent loads per record)
            for (int c = 0; c < L; c += 8) {
                const int b0 = (odd + c) >> 1;
                u64 win = 0;
#pragma unroll
                for (int j = 0; j < 5; ++j)
                    if (2 * (b0 + j) < odd + L) win |= (u64)sp[b0 + j] << (8 * j);
                win >>= 4 * odd;
                int32_t nd[8], nl[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {                    // the record's nibble: step + 8, 0 = listed
                    const int nib = (int)(win >> (4 * k)) & 15;
                    if (c + k < L && c + k > 0) node = nib ? node + nib - 8 : wide_node_of(w, lo + c + k);
                    nd[k] = node;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) nl[k] = c + k < L ? node_len[nd[k]] : 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = c + k;
                    if (i < L) {
                        // first / last record: the shipped byte; in between the node is covered completely.  Records with
                        // an explicit coverage (byte 255, or listed) are overwritten by k_apply_exceptions.
                        const int32_t al = i == 0 ? al_first : (i == L - 1 ? al_last : (nl[k] & 0xFFFF));
                        if (staged) {
                            s_node[base + i] = nd[k];
                            s_alen[base + i] = al | (nl[k] << 16);
                        } else {
                            walk_node[lo + i] = nd[k];
                            walk_alen[lo + i] = al | (nl[k] << 16);
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (staged) {
            const int n = (int)(r1 - r0);
            for (int i = threadIdx.x; i < n; i += kWalkTile) {
                walk_node[r0 + i] = s_node[i];
                walk_alen[r0 + i] = s_alen[i];
            }
        }
        __syncthreads();
    }
}
__global__ void k_apply_exceptions(const i64* exc_idx, i64 n_exc, int32_t* walk_alen) {
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < n_exc; k += (i64)gridDim.x * blockDim.x)
        walk_alen[exc_idx[k]] = -(int32_t)(k + 1);
}

static int scan_u8(const uint8_t* cnt, i64 n, i64* off, i64* scratch, cudaStream_t st) {
    if (n == 0) {
        cudaMemsetAsync(off, 0, sizeof(i64), st);
        return SFB_OK;
    }
    const i64 tiles = (n + kScanTile - 1) / kScanTile;
    k_scan_partials<<<(unsigned)tiles, kScanBlock, 0, st>>>(cnt, n, scratch);
    k_scan_tiles<<<1, kScanBlock, 0, st>>>(scratch, tiles);
    k_scan_write<<<(unsigned)tiles, kScanBlock, 0, st>>>(cnt, n, scratch, off);
    return check_launch("scan_u8");
}

int launch_expand(const sfb_wire* w, const int32_t* node_len, const sfb_reads* dst, i64* scratch, cudaStream_t st) {
    i64* grp_aln_off = const_cast<i64*>((const i64*)dst->grp_aln_off);
    i64* aln_walk_off = const_cast<i64*>((const i64*)dst->aln_walk_off);
    int32_t* mate_seq_len = const_cast<int32_t*>(dst->mate_seq_len);
    uint8_t* aln_bins = const_cast<uint8_t*>(dst->aln_bins);
    int32_t* walk_node = const_cast<int32_t*>(dst->walk_node);
    int32_t* walk_alen = const_cast<int32_t*>(dst->walk_alen);
    SFB_TRY(scan_u8(w->mate_naln, 2 * w->n_groups, grp_aln_off, scratch, st));
    SFB_TRY(scan_u8(w->aln_len, w->n_alns, aln_walk_off, scratch, st));
    if (w->n_groups) k_expand_seq_len<<<kSMs * 4, 256, 0, st>>>(*w, mate_seq_len);
    if (w->n_seq_exc) k_apply_seq_exceptions<<<blocks_for(w->n_seq_exc, 256), 256, 0, st>>>(*w, mate_seq_len);
    if (w->n_alns) k_expand_bins<<<kSMs * 8, 256, 0, st>>>(*w, aln_bins);
    if (w->n_bins_exc) k_apply_bins_exceptions<<<blocks_for(w->n_bins_exc, 256), 256, 0, st>>>(*w, aln_bins);
    if (w->n_records) {
        static OccCache cache;               // one wave of resident blocks (the block loop strides over the tiles of 256 walks)
        const unsigned grid = persistent_grid(k_expand_walks, kWalkTile, blocks_for(w->n_alns, kWalkTile), cache);
        k_expand_walks<<<grid, kWalkTile, 0, st>>>(*w, aln_walk_off, node_len, walk_node, walk_alen);
    }
    if (w->n_exc) k_apply_exceptions<<<blocks_for(w->n_exc, 256), 256, 0, st>>>((const i64*)w->exc_idx, w->n_exc, walk_alen);
    return check_launch("expand_reads");
}

}  // namespace sfb
