// libsfb200: hand-written sm_100a kernels for StrainFLAIR's query-side abundance pass.
// Public surface: include/sfb200.h.  Build: strainflair_b200/build.py (nvcc -gencode arch=compute_100a,code=sm_100a).
//
// All kernels are bandwidth / atomic bound integer+fp64 work (no tensor cores, by design):
//   k_match          one thread per retained alignment, walk vs. path-slot compares through the
//                    node->occurrence index (L2-resident gathers), warp-aggregated claim of output space
//   k_classify       one thread per read group, the pass-1 decision tree on the match lists
//   k_pass1_scatter  record-parallel: coalesced walk_alen stream, fp64 RED into the slot array
//   k_pass2          deferred groups, proportional redistribution
//   k_path_reduce    one warp per colored path, fixed-order fp64 tree
//   k_strain_*       compute_strains_abundance on the gene table
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

// =============================================================================================
// match: Pangenome.get_matching_path (json2csv.py:341-363) for the walk and the reversed walk
// =============================================================================================
constexpr int kMatchLocal = 8;   // matches kept in registers before spilling to a second sweep

struct MatchStats {
    int cand, cmp;
};

template <bool kEmit>
__device__ __forceinline__ int match_sweep(const sfb_graph& g, const sfb_reads& r, i64 w0, int L, uint32_t* keep,
                                           int32_t* out, MatchStats& st) {
    int n = 0;
    const int first = r.walk_node[w0];
    // forward: occurrences of walk[0], then compare walk[1..]
    for (int o = g.occ_off[first], oe = g.occ_off[first + 1]; o < oe; ++o) {
        const int s = g.occ_slot[o];
        bool ok = true;
        int i = 1;
        for (; i < L; ++i)
            if (g.slot_node[s + i] != r.walk_node[w0 + i]) { ok = false; break; }   // sentinel (-1) stops overruns
        st.cand += 1;
        st.cmp += min(i, L - 1);
        if (ok) {
            if (kEmit) out[n] = s; else if (n < kMatchLocal) keep[n] = (uint32_t)s;
            ++n;
        }
    }
    if (L > 1) {   // reversed walk (json2csv.py:861-866): occurrences of walk[L-1]
        const int last = r.walk_node[w0 + L - 1];
        for (int o = g.occ_off[last], oe = g.occ_off[last + 1]; o < oe; ++o) {
            const int s = g.occ_slot[o];
            bool ok = true;
            int i = 1;
            for (; i < L; ++i)
                if (g.slot_node[s + i] != r.walk_node[w0 + L - 1 - i]) { ok = false; break; }
            st.cand += 1;
            st.cmp += min(i, L - 1);
            if (ok) {
                const uint32_t code = (uint32_t)s | SFB_MATCH_REV;
                if (kEmit) out[n] = (int32_t)code; else if (n < kMatchLocal) keep[n] = code;
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
    const i64 n_round = (r.n_alns + 255) & ~(i64)255;
    for (i64 a = (i64)blockIdx.x * blockDim.x + threadIdx.x; a < n_round; a += (i64)gridDim.x * blockDim.x) {
        const bool live = a < r.n_alns;
        uint32_t keep[kMatchLocal];
        int n = 0, L = 0;
        i64 w0 = 0;
        if (live) {
            w0 = r.aln_walk_off[a];
            L = (int)(r.aln_walk_off[a + 1] - w0);
            if (L > 0) n = match_sweep<false>(g, r, w0, L, keep, nullptr, st);
            tuples += n;
        }
        // warp-aggregated claim of match_buf space: [count, codes...] padded to an even number of ints
        const unsigned need = n > 1 ? (unsigned)((n + 2) & ~1) : 0u;
        unsigned incl = need;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane_id() >= d) incl += t;
        }
        const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
        u64 base = 0;
        if (total) {
            if (lane_id() == 31) base = atomicAdd(w.cursor, (u64)total);
            base = __shfl_sync(0xffffffffu, base, 31);
        }
        if (!live) continue;
        if (n == 0) {
            w.aln_match[a] = SFB_MATCH_NONE;
        } else if (n == 1) {
            w.aln_match[a] = keep[0];
        } else {
            const u64 pos = base + (incl - need);
            if (pos + need <= (u64)w.match_cap) {
                int32_t* out = w.match_buf + pos;
                out[0] = n;
                if (n <= kMatchLocal) {
                    for (int k = 0; k < n; ++k) out[1 + k] = (int32_t)keep[k];
                } else {
                    MatchStats scratch{0, 0};
                    match_sweep<true>(g, r, w0, L, nullptr, out + 1, scratch);
                }
                w.aln_match[a] = SFB_MATCH_MULTI | (uint32_t)(pos >> 1);
            } else {
                w.aln_match[a] = SFB_MATCH_NONE;
                SFB_COUNT(sc, SFB_C_MATCH_OVERFLOW, (int)need);
            }
        }
    }
    // statistics: warp tree, then one shared-memory add per warp, one global add per block
    int c0 = st.cand, c1 = st.cmp, c2 = tuples;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, d);
        c1 += __shfl_xor_sync(0xffffffffu, c1, d);
        c2 += __shfl_xor_sync(0xffffffffu, c2, d);
    }
    if (lane_id() == 0) {
        SFB_COUNT(sc, SFB_C_ST_CAND, c0);
        SFB_COUNT(sc, SFB_C_ST_COMPARES, c1);
        SFB_COUNT(sc, SFB_C_ST_TUPLES, c2);
    }
    __syncthreads();
    block_counters_flush(sc, counters);
}

// =============================================================================================
// classify: the pass-1 decision tree (json2csv.py:828-1053)
// =============================================================================================
struct Dev {
    sfb_graph g;
    sfb_reads r;
    sfb_work w;
    sfb_accum acc;
};

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
    atomicAdd(&d.acc.uniq_norm[pid], (double)seq_len / (double)d.g.path_seq_len[pid]);
    hist_add(d, sc, pid, a);
}

// cluster of the first path through the walk's first node (json2csv.py:875,901,1036); -1 = None, -2 = no path
__device__ __forceinline__ int first_cluster(const Dev& d, i64 a) {
    const int n0 = d.r.walk_node[d.r.aln_walk_off[a]];
    const int o = d.g.occ_off[n0];
    if (o == d.g.occ_off[n0 + 1]) return -2;
    return d.g.path_cluster[d.g.slot_path[d.g.occ_slot[o]]];
}
__device__ __forceinline__ void book(const Dev& d, BlockCounters& sc, const Mate& mt, int cluster) {
    if (cluster < 0) { SFB_COUNT(sc, SFB_C_NO_CLUSTER, 1); return; }
    d.w.aln_assign[mt.a0] = -(cluster + 2);
}

// Strain reduction (json2csv.py:964-986): number of entries of the reduced list of mate `mt`
// (a tuple counts once per common strain it carries) and whether any common strain exists.
struct StrainCut {
    bool any;
    int n_new[2];
};
__device__ __forceinline__ StrainCut strain_cut(const Dev& d, const Mate* mt) {
    StrainCut c;
    c.any = false;
    c.n_new[0] = c.n_new[1] = 0;
    for (int word = 0; word < (int)d.g.mask_words; ++word) {
        const u64 common = mate_strain_union(d.g, d.w, mt[0], word) & mate_strain_union(d.g, d.w, mt[1], word);
        if (!common) continue;
        c.any = true;
        for (int k = 0; k < 2; ++k)
            for_tuples(d.g, d.w, mt[k], [&](int, uint32_t, int p, int) {
                c.n_new[k] += __popcll(d.g.path_smask[(size_t)p * d.g.mask_words + word] & common);
                return true;
            });
    }
    return c;
}
// multiplicity of path p in the reduced list
__device__ __forceinline__ int strain_mult(const Dev& d, const Mate* mt, int p) {
    int m = 0;
    for (int word = 0; word < (int)d.g.mask_words; ++word) {
        const u64 common = mate_strain_union(d.g, d.w, mt[0], word) & mate_strain_union(d.g, d.w, mt[1], word);
        m += __popcll(d.g.path_smask[(size_t)p * d.g.mask_words + word] & common);
    }
    return m;
}

// single-end style (json2csv.py:1003-1050)
__device__ __forceinline__ int classify_single(const Dev& d, BlockCounters& sc, Mate& mt, bool& defer) {
    SFB_COUNT(sc, SFB_C_SINGLES, 1);
    if (mt.na == 0) { SFB_COUNT(sc, SFB_C_FILTERED, 1); return SFB_FILTERED; }
    if (mt.na > 1) { SFB_COUNT(sc, SFB_C_SECOND_PASS, 1); defer = true; return SFB_DEFER; }
    const int n = match_count(d.w, mt.a0);
    if (n > 1) { SFB_COUNT(sc, SFB_C_SECOND_PASS, 1); defer = true; return SFB_DEFER; }
    if (n == 0) {
        SFB_COUNT(sc, SFB_C_UNASSIGNED, 1);
        const int c = first_cluster(d, mt.a0);
        if (c == -2) SFB_COUNT(sc, SFB_C_NO_PATH, 1); else book(d, sc, mt, c);
        return SFB_UNASSIGNED_BOOK;
    }
    SFB_COUNT(sc, SFB_C_ASSIGNED_U, 1);      // counted, never accumulated (json2csv.py:1047-1053, Q1)
    SFB_COUNT(sc, SFB_C_SE, 1);
    return SFB_SE_COUNTED;
}

__global__ void __launch_bounds__(128) k_classify(Dev d) {
    __shared__ BlockCounters sc;
    block_counters_zero(sc);
    __syncthreads();
    const i64 n_round = (d.r.n_groups + 127) & ~(i64)127;
    for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < n_round; g += (i64)gridDim.x * blockDim.x) {
    bool defer = false;
    if (g < d.r.n_groups) {
        const int nm = d.r.grp_nmates[g];
        Mate mt[2];
        mate_load(d.r, d.w, g, 0, mt[0]);
        mate_load(d.r, d.w, g, 1, mt[1]);
        for (int k = 0; k < 2; ++k)
            for (int j = 0; j < mt[k].na; ++j) d.w.aln_assign[mt[k].a0 + j] = -1;
        int code[2] = {SFB_ABSENT, SFB_ABSENT};
        SFB_COUNT(sc, SFB_C_NB_READS, nm);
        int single = nm == 1 ? 0 : -1;       // mate index handled single-end style, -1 = none
        if (nm == 2) {
            SFB_COUNT(sc, SFB_C_PAIRS, 1);
            if (mt[0].na == 0 && mt[1].na == 0) {
                SFB_COUNT(sc, SFB_C_FILTERED, 2);
                code[0] = code[1] = SFB_FILTERED;
            } else if (mt[0].na == 0 || mt[1].na == 0) {
                SFB_COUNT(sc, SFB_C_FILTERED, 1);
                SFB_COUNT(sc, SFB_C_ONE_NOALIGN, 1);
                const int gone = mt[0].na == 0 ? 0 : 1;
                code[gone] = SFB_FILTERED;
                single = 1 - gone;
            } else {
                mate_count_tuples(d.w, mt[0]);
                mate_count_tuples(d.w, mt[1]);
                if (mt[0].nt == 0 && mt[1].nt == 0) {
                    // both unassigned (json2csv.py:868-891)
                    SFB_COUNT(sc, SFB_C_UNASSIGNED, 2);
                    int n_common = 0, common = 0;
                    bool bad = false;
                    for (int i = 0; i < mt[0].na; ++i) {
                        const int ci = first_cluster(d, mt[0].a0 + i);
                        if (ci == -2) bad = true;
                        bool seen = false;
                        for (int j = 0; j < i; ++j) seen |= first_cluster(d, mt[0].a0 + j) == ci;
                        if (seen) continue;
                        bool other = false;
                        for (int j = 0; j < mt[1].na; ++j) other |= first_cluster(d, mt[1].a0 + j) == ci;
                        if (other) { ++n_common; common = ci; }
                    }
                    for (int j = 0; j < mt[1].na; ++j) bad |= first_cluster(d, mt[1].a0 + j) == -2;
                    code[0] = code[1] = SFB_UNASSIGNED;
                    if (bad) {
                        SFB_COUNT(sc, SFB_C_NO_PATH, 1);
                    } else if (n_common <= 1) {
                        for (int k = 0; k < 2; ++k) {
                            int len = mt[k].na, cl = first_cluster(d, mt[k].a0);
                            if (n_common == 1) {
                                len = 0;
                                for (int j = 0; j < mt[k].na; ++j) len += first_cluster(d, mt[k].a0 + j) == common;
                                cl = common;
                            }
                            if (len == 1) { book(d, sc, mt[k], cl); code[k] = SFB_UNASSIGNED_BOOK; }
                        }
                    }
                } else if (mt[0].nt == 0 || mt[1].nt == 0) {
                    // one mate without colored path (json2csv.py:895-916)
                    SFB_COUNT(sc, SFB_C_UNASSIGNED, 1);
                    SFB_COUNT(sc, SFB_C_ONE_NOCP, 1);
                    const int gone = mt[0].nt == 0 ? 0 : 1;
                    code[gone] = SFB_UNASSIGNED;
                    bool bad = false;
                    for (int j = 0; j < mt[gone].na; ++j) bad |= first_cluster(d, mt[gone].a0 + j) == -2;
                    if (bad) {
                        SFB_COUNT(sc, SFB_C_NO_PATH, 1);
                    } else if (mt[gone].na == 1) {
                        book(d, sc, mt[gone], first_cluster(d, mt[gone].a0));
                        code[gone] = SFB_UNASSIGNED_BOOK;
                    }
                    single = 1 - gone;
                } else {
                    // both mates have colored paths: intersect path ids (json2csv.py:922-959)
                    int n_inter = 0, p_inter = -1;
                    for_tuples(d.g, d.w, mt[0], [&](int i, uint32_t, int p, int) {
                        if (first_of_pid(d.g, d.w, mt[0], i, p) && count_pid(d.g, d.w, mt[1], p) > 0) {
                            if (n_inter++ == 0) p_inter = p;
                        }
                        return true;
                    });
                    if (n_inter > 1) {
                        SFB_COUNT(sc, SFB_C_SECOND_PASS, 2);
                        defer = true;
                        code[0] = code[1] = SFB_DEFER;
                    } else if (n_inter == 1) {
                        for (int k = 0; k < 2; ++k) {
                            if (count_pid(d.g, d.w, mt[k], p_inter) > 1) {
                                SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
                                code[k] = SFB_MULTI_ALIGN;
                                continue;
                            }
                            for_tuples(d.g, d.w, mt[k], [&](int, uint32_t c, int p, int aloc) {
                                if (p != p_inter) return true;
                                assign_unique(d, sc, mt[k].a0 + aloc, c, p, mt[k].seq_len);
                                return false;
                            });
                            SFB_COUNT(sc, SFB_C_ASSIGNED_U, 1);
                            SFB_COUNT(sc, SFB_C_SAME_CP, 1);
                            if (mt[k].nt > 1) SFB_COUNT(sc, SFB_C_AMBIG_CP_RESOLVED, 1);
                            code[k] = SFB_ASSIGN_SAME;
                        }
                    } else {
                        // no common path: intersect strains (json2csv.py:961-1000)
                        const StrainCut cut = strain_cut(d, mt);
                        if (!cut.any) {
                            SFB_COUNT(sc, SFB_C_INCONSIST, 2);
                            code[0] = code[1] = SFB_INCONSIST;
                        } else {
                            for (int k = 0; k < 2; ++k) {
                                const int nn = cut.n_new[k];
                                if (nn < mt[k].nt) SFB_COUNT(sc, nn == 1 ? SFB_C_AMBIG_RESOLVED : SFB_C_AMBIG_REDUCED, 1);
                                if (nn > 1) {
                                    SFB_COUNT(sc, SFB_C_SECOND_PASS, 1);
                                    defer = true;
                                    code[k] = SFB_DEFER;
                                    continue;
                                }
                                // exactly one tuple survives, once
                                for_tuples(d.g, d.w, mt[k], [&](int, uint32_t c, int p, int aloc) {
                                    if (strain_mult(d, mt, p) == 0) return true;
                                    if (count_pid(d.g, d.w, mt[k], p) == 1) {
                                        assign_unique(d, sc, mt[k].a0 + aloc, c, p, mt[k].seq_len);
                                        SFB_COUNT(sc, SFB_C_ASSIGNED_U, 1);
                                        SFB_COUNT(sc, SFB_C_DIFF_CP, 1);
                                        code[k] = SFB_ASSIGN_DIFF;
                                    } else {
                                        SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
                                        code[k] = SFB_MULTI_ALIGN;
                                    }
                                    return false;
                                });
                            }
                        }
                    }
                }
            }
        }
        if (single >= 0) code[single] = classify_single(d, sc, mt[single], defer);
        d.w.grp_code[g] = (uint8_t)(code[0] | (code[1] << 4));
    }
    // warp-aggregated append to the deferred list
    const unsigned ballot = __ballot_sync(0xffffffffu, defer);
    if (ballot) {
        u64 base = 0;
        const int leader = __ffs(ballot) - 1;
        if (lane_id() == leader) base = atomicAdd(d.w.cursor + 1, (u64)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (defer) d.w.deferred[base + __popc(ballot & ((1u << lane_id()) - 1))] = (int32_t)g;
    }
    }
    __syncthreads();
    block_counters_flush(sc, d.acc.counters);
}

// =============================================================================================
// pass 1 scatter: uniq_abund[slot+i] += cov[i] (json2csv.py:329-330), record-parallel
// =============================================================================================
constexpr int kTileAlns = 256;

__global__ void __launch_bounds__(256) k_pass1_scatter(sfb_graph g, sfb_reads r, sfb_work w, double* uniq_abund,
                                                       long long* counters) {
    __shared__ int s_off[kTileAlns + 1];
    __shared__ int32_t s_code[kTileAlns];
    __shared__ int s_done;
    if (threadIdx.x == 0) s_done = 0;
    __syncthreads();
    int done = 0;
    for (i64 tile = blockIdx.x; tile * kTileAlns < r.n_alns; tile += gridDim.x) {
    const i64 a_lo = tile * kTileAlns;
    const int na = (int)min((i64)kTileAlns, r.n_alns - a_lo);
    const i64 r_lo = r.aln_walk_off[a_lo];
    for (int j = threadIdx.x; j <= na; j += blockDim.x) s_off[j] = (int)(r.aln_walk_off[a_lo + j] - r_lo);
    for (int j = threadIdx.x; j < na; j += blockDim.x) s_code[j] = w.aln_assign[a_lo + j];
    __syncthreads();
    const int nrec = s_off[na];
    for (int k = threadIdx.x; k < nrec; k += blockDim.x) {
        int lo = 0, hi = na;                    // largest j with s_off[j] <= k
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_off[mid] <= k) lo = mid; else hi = mid;
        }
        const int32_t code = s_code[lo];
        if (code < 0) continue;
        const int i = k - s_off[lo];
        const int L = s_off[lo + 1] - s_off[lo];
        const double cov = record_cov(g, r, r_lo + k, (uint32_t)code, i, L);
        atomicAdd(&uniq_abund[((uint32_t)code & SFB_SLOT_MASK) + i], cov);
        ++done;
    }
    __syncthreads();
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) done += __shfl_xor_sync(0xffffffffu, done, d);
    if (lane_id() == 0 && done) atomicAdd(&s_done, done);
    __syncthreads();
    if (threadIdx.x == 0 && s_done) atomicAdd((u64*)&counters[SFB_C_ST_P1_RECORDS], (u64)s_done);
}

// =============================================================================================
// pass 2: proportional redistribution (json2csv.py:1085-1274)
// =============================================================================================
// json2csv.py:1155-1158 (and :1207-1210, :1265-1268); `times` > 1 when the strain reduction
// duplicated the candidate (Q4): the reference adds the same terms `times` times.
__device__ __forceinline__ void apply_multi(const Dev& d, BlockCounters& sc, i64 a, uint32_t code, int pid,
                                            int seq_len, double ratio, int times) {
    const double tr = ratio * (double)times;
    atomicAdd(&d.acc.mult_reads[pid], tr);
    atomicAdd(&d.acc.mult_norm[pid], (ratio * (double)seq_len / (double)d.g.path_seq_len[pid]) * (double)times);
    atomicAdd(&d.acc.mult_hits[pid], (unsigned)times);
    const i64 w0 = d.r.aln_walk_off[a];
    const int L = (int)(d.r.aln_walk_off[a + 1] - w0);
    const uint32_t slot0 = code & SFB_SLOT_MASK;
    for (int i = 0; i < L; ++i)
        atomicAdd(&d.acc.mult_abund[slot0 + i], (record_cov(d.g, d.r, w0 + i, code, i, L) * ratio) * (double)times);
    SFB_COUNT(sc, SFB_C_ST_P2_APPLIED, 1);
    SFB_COUNT(sc, SFB_C_ST_P2_RECORDS, L);
}

__device__ __forceinline__ double ratio_of(long long u, long long total, int n) {
    return total == 0 ? 1.0 / (double)n : (double)u / (double)total;
}

__device__ __forceinline__ int pass2_single(const Dev& d, BlockCounters& sc, const Mate& mt, const long long* U) {
    bool hit = false;
    for (int j = 0; j < mt.na; ++j) {
        const i64 a = mt.a0 + j;
        const int n = match_count(d.w, a);
        if (n == 0) continue;
        hit = true;
        long long total = 0;
        for (int k = 0; k < n; ++k) {
            const int p = d.g.slot_path[match_get(d.w, a, k) & SFB_SLOT_MASK];
            total += U[p];
            hist_add(d, sc, p, a);
        }
        for (int k = 0; k < n; ++k) {
            const uint32_t c = match_get(d.w, a, k);
            const int p = d.g.slot_path[c & SFB_SLOT_MASK];
            apply_multi(d, sc, a, c, p, mt.seq_len, ratio_of(U[p], total, n), 1);
        }
    }
    if (hit) {
        SFB_COUNT(sc, SFB_C_ASSIGNED_M, 1);
        SFB_COUNT(sc, SFB_C_SE_M, 1);
        return SFB_P2_SE_ASSIGNED;
    }
    SFB_COUNT(sc, SFB_C_UNASSIGNED, 1);
    return SFB_P2_UNASSIGNED;
}

__global__ void __launch_bounds__(128) k_pass2(Dev d, const long long* __restrict__ U) {
    __shared__ BlockCounters sc;
    block_counters_zero(sc);
    __syncthreads();
    const i64 n_def = (i64)d.w.cursor[1];
    for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < n_def; q += (i64)gridDim.x * blockDim.x) {
        const i64 g = d.w.deferred[q];
        const int nm = d.r.grp_nmates[g];
        Mate mt[2];
        mate_load(d.r, d.w, g, 0, mt[0]);
        mate_load(d.r, d.w, g, 1, mt[1]);
        int code[2] = {SFB_P2_NONE, SFB_P2_NONE};
        int single = nm == 1 ? 0 : -1;
        if (nm == 2) {
            if (mt[0].na == 0 || mt[1].na == 0) {
                single = mt[0].na == 0 ? 1 : 0;
            } else {
                mate_count_tuples(d.w, mt[0]);
                mate_count_tuples(d.w, mt[1]);
                if (mt[0].nt == 0 || mt[1].nt == 0) {
                    single = mt[0].nt == 0 ? 1 : 0;
                } else {
                    // intersection of path ids: size and sum of unique counts (json2csv.py:1127-1141)
                    int n_inter = 0;
                    long long total = 0;
                    for_tuples(d.g, d.w, mt[0], [&](int i, uint32_t, int p, int) {
                        if (first_of_pid(d.g, d.w, mt[0], i, p) && count_pid(d.g, d.w, mt[1], p) > 0) {
                            ++n_inter;
                            total += U[p];
                        }
                        return true;
                    });
                    if (n_inter >= 1) {
                        for (int k = 0; k < 2; ++k) {
                            bool simple = true;   // every common path has exactly one tuple in this mate
                            for_tuples(d.g, d.w, mt[k], [&](int i, uint32_t, int p, int) {
                                if (count_pid(d.g, d.w, mt[1 - k], p) > 0 && count_pid(d.g, d.w, mt[k], p) != 1) {
                                    simple = false;
                                    return false;
                                }
                                return true;
                            });
                            if (!simple) {
                                SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
                                code[k] = SFB_P2_MULTI_ALIGN;
                                continue;
                            }
                            SFB_COUNT(sc, SFB_C_ASSIGNED_M, 1);
                            code[k] = SFB_P2_ASSIGNED;
                            for_tuples(d.g, d.w, mt[k], [&](int, uint32_t c, int p, int aloc) {
                                if (count_pid(d.g, d.w, mt[1 - k], p) > 0)
                                    apply_multi(d, sc, mt[k].a0 + aloc, c, p, mt[k].seq_len, ratio_of(U[p], total, n_inter), 1);
                                return true;
                            });
                        }
                    } else {
                        const StrainCut cut = strain_cut(d, mt);
                        if (!cut.any) {
                            SFB_COUNT(sc, SFB_C_INCONSIST_M, (mt[0].nt == 1 || mt[1].nt == 1) ? 1 : 2);
                            code[0] = code[1] = SFB_P2_INCONSIST;
                        } else {
                            for (int k = 0; k < 2; ++k) {
                                const int nn = cut.n_new[k];
                                if (nn <= 1) continue;      // handled in pass 1 (json2csv.py:1185)
                                bool simple = true;
                                long long tot = 0;
                                for_tuples(d.g, d.w, mt[k], [&](int, uint32_t, int p, int) {
                                    const int times = strain_mult(d, mt, p);
                                    if (times == 0) return true;
                                    if (count_pid(d.g, d.w, mt[k], p) != 1) { simple = false; return false; }
                                    tot += U[p] * times;
                                    return true;
                                });
                                if (!simple) {
                                    SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
                                    code[k] = SFB_P2_MULTI_ALIGN;
                                    continue;
                                }
                                SFB_COUNT(sc, SFB_C_ASSIGNED_M, 1);
                                code[k] = SFB_P2_ASSIGNED;
                                for_tuples(d.g, d.w, mt[k], [&](int, uint32_t c, int p, int aloc) {
                                    const int times = strain_mult(d, mt, p);
                                    if (times)
                                        apply_multi(d, sc, mt[k].a0 + aloc, c, p, mt[k].seq_len, ratio_of(U[p], tot, nn), times);
                                    return true;
                                });
                            }
                        }
                    }
                }
            }
        }
        if (single >= 0) code[single] = pass2_single(d, sc, mt[single], U);
        d.w.grp_code2[g] = (uint8_t)(code[0] | (code[1] << 4));
    }
    __syncthreads();
    block_counters_flush(sc, d.acc.counters);
}

// =============================================================================================
// path reduce: print_gene_table's per-path statistics (json2csv.py:440-469), one warp per path
// =============================================================================================
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__global__ void __launch_bounds__(256) k_path_reduce(sfb_graph g, sfb_accum acc, i64 p_begin, i64 p_end, double* table) {
    const i64 p = p_begin + ((i64)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
    if (p >= p_end) return;
    const int lo = g.path_off[p], hi = g.path_off[p + 1] - 1;
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
        t[0] = acc.uniq_norm[p];
        t[1] = (double)acc.uniq_reads[p] + acc.mult_reads[p];
        t[2] = acc.uniq_norm[p] + acc.mult_norm[p];
        t[3] = su / n;
        t[4] = cu ? su / (double)cu : nan;
        t[5] = sum / n;
        t[6] = cum ? sum / (double)cum : nan;
        t[7] = (double)cc / n;
    }
}

// =============================================================================================
// cluster level: strain copy counts per cluster (json2csv.py:544)
// =============================================================================================
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

// =============================================================================================
// strain level: compute_strains_abundance.py:52-69
// =============================================================================================
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
// k-th smallest (0-based) of v[0..n) by MSB-first radix selection; whole block participates
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

}  // namespace sfb

// =============================================================================================
// C ABI
// =============================================================================================
using namespace sfb;

static int check_graph(const sfb_graph* g) {
    SFB_REQUIRE(g, "graph is null");
    SFB_REQUIRE(g->n_nodes >= 0 && g->n_paths >= 0 && g->n_slots >= 0 && g->n_strains >= 0, "negative size");
    SFB_REQUIRE(g->n_slots <= SFB_MAX_SLOTS, "n_slots exceeds 2^30-1");
    SFB_REQUIRE(g->mask_words * 64 >= g->n_strains, "mask_words too small");
    SFB_REQUIRE(g->slot_node && g->slot_len && g->slot_path && g->path_off && g->occ_off && g->occ_slot &&
                    g->path_seq_len && g->path_nstrain && g->path_strain0 && g->path_smask && g->path_cluster,
                "null graph array");
    return SFB_OK;
}
static int check_reads(const sfb_reads* r) {
    SFB_REQUIRE(r, "reads is null");
    SFB_REQUIRE(r->n_groups >= 0 && r->n_alns >= 0 && r->n_records >= 0 && r->n_exc >= 0, "negative size");
    SFB_REQUIRE(r->n_groups < 0x7fffffffLL, "more than 2^31-1 groups in one shard");
    SFB_REQUIRE(r->grp_aln_off && r->aln_walk_off, "null reads array");
    SFB_REQUIRE(r->n_groups == 0 || (r->grp_nmates && r->mate_seq_len), "null reads array");
    SFB_REQUIRE(r->n_alns == 0 || (r->aln_bins && r->walk_node && r->walk_alen), "null reads array");
    SFB_REQUIRE(r->n_exc == 0 || r->exc_cov, "exc_cov is null");
    return SFB_OK;
}
static int check_work(const sfb_work* w) {
    SFB_REQUIRE(w, "work is null");
    SFB_REQUIRE(w->aln_match && w->cursor && w->grp_code && w->grp_code2 && w->aln_assign && w->deferred,
                "null work array");
    SFB_REQUIRE(w->match_cap >= 0 && w->match_cap < 0x100000000LL && (w->match_cap == 0 || w->match_buf),
                "bad match buffer");
    return SFB_OK;
}
static int check_accum(const sfb_accum* a) {
    SFB_REQUIRE(a, "accum is null");
    SFB_REQUIRE(a->uniq_reads && a->uniq_norm && a->mult_reads && a->mult_norm && a->mult_hits && a->uniq_abund &&
                    a->mult_abund && a->hist && a->counters,
                "null accumulator");
    return SFB_OK;
}
#define SFB_TRY(x)                 \
    do {                           \
        int rc_ = (x);             \
        if (rc_ != SFB_OK) return rc_; \
    } while (0)

static inline unsigned blocks_for(i64 n, int per) { return (unsigned)((n + per - 1) / per); }
constexpr int kSMs = 148;   // B200: persistent grids are sized in multiples of the SM count

extern "C" {

int sfb_abi_version(void) { return SFB_ABI_VERSION; }
const char* sfb_version(void) { return "sfb200 0.1 (sm_100a)"; }
const char* sfb_last_error(void) { return g_err; }

int sfb_match(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, void* stream) {
    SFB_TRY(check_graph(g));
    SFB_TRY(check_reads(r));
    SFB_TRY(check_work(w));
    SFB_TRY(check_accum(acc));
    if (r->n_alns == 0) return SFB_OK;
    const unsigned grid = (unsigned)min((i64)kSMs * 8, (i64)blocks_for(r->n_alns, 256));
    k_match<<<grid, 256, 0, (cudaStream_t)stream>>>(*g, *r, *w, acc->counters);
    return check_launch("k_match");
}

int sfb_classify(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, void* stream) {
    SFB_TRY(check_graph(g));
    SFB_TRY(check_reads(r));
    SFB_TRY(check_work(w));
    SFB_TRY(check_accum(acc));
    if (r->n_groups == 0) return SFB_OK;
    Dev d{*g, *r, *w, *acc};
    const unsigned grid = (unsigned)min((i64)kSMs * 16, (i64)blocks_for(r->n_groups, 128));
    k_classify<<<grid, 128, 0, (cudaStream_t)stream>>>(d);
    return check_launch("k_classify");
}

int sfb_pass1_scatter(const sfb_graph* g, const sfb_reads* r, const sfb_work* w, sfb_accum* acc, void* stream) {
    SFB_TRY(check_graph(g));
    SFB_TRY(check_reads(r));
    SFB_TRY(check_work(w));
    SFB_TRY(check_accum(acc));
    if (r->n_alns == 0) return SFB_OK;
    const unsigned grid = (unsigned)min((i64)kSMs * 8, (i64)blocks_for(r->n_alns, kTileAlns));
    k_pass1_scatter<<<grid, 256, 0, (cudaStream_t)stream>>>(*g, *r, *w, acc->uniq_abund, acc->counters);
    return check_launch("k_pass1_scatter");
}

int sfb_pass2(const sfb_graph* g, const sfb_reads* r, sfb_work* w, sfb_accum* acc, const long long* uniq_reads_global,
              void* stream) {
    SFB_TRY(check_graph(g));
    SFB_TRY(check_reads(r));
    SFB_TRY(check_work(w));
    SFB_TRY(check_accum(acc));
    SFB_REQUIRE(uniq_reads_global, "uniq_reads_global is null");
    if (r->n_groups == 0) return SFB_OK;
    Dev d{*g, *r, *w, *acc};
    const unsigned grid = (unsigned)min((i64)kSMs * 16, (i64)blocks_for(r->n_groups, 128));
    k_pass2<<<grid, 128, 0, (cudaStream_t)stream>>>(d, uniq_reads_global);
    return check_launch("k_pass2");
}

int sfb_path_reduce(const sfb_graph* g, const sfb_accum* acc, int64_t p_begin, int64_t p_end, double* table,
                    void* stream) {
    SFB_TRY(check_graph(g));
    SFB_TRY(check_accum(acc));
    SFB_REQUIRE(0 <= p_begin && p_begin <= p_end && p_end <= g->n_paths, "bad path range");
    SFB_REQUIRE(table || p_begin == p_end, "table is null");
    if (p_begin == p_end) return SFB_OK;
    k_path_reduce<<<blocks_for((p_end - p_begin) * kWarp, 256), 256, 0, (cudaStream_t)stream>>>(*g, *acc, p_begin, p_end,
                                                                                                table);
    return check_launch("k_path_reduce");
}

int sfb_cluster_strain_counts(int64_t n_clusters, int64_t n_strains, const int32_t* clu_off, const int32_t* clu_paths,
                              const int32_t* pstr_off, const int32_t* pstr_id, const int32_t* pstr_cnt, long long* out,
                              void* stream) {
    SFB_REQUIRE(n_clusters >= 0 && n_strains >= 0, "negative size");
    if (n_clusters == 0 || n_strains == 0) return SFB_OK;
    SFB_REQUIRE(clu_off && clu_paths && pstr_off && pstr_id && pstr_cnt && out, "null array");
    k_cluster_counts<<<blocks_for(n_clusters, 128), 128, 0, (cudaStream_t)stream>>>(n_clusters, n_strains, clu_off,
                                                                                    clu_paths, pstr_off, pstr_id,
                                                                                    pstr_cnt, out);
    return check_launch("k_cluster_counts");
}

int sfb_strain_aggregate(int64_t n_rows, int64_t n_strains, const int32_t* row_off, const int32_t* row_strain,
                         const double* row_count, const double* mam, const double* mam_nz, const double* ratio_cov,
                         double thr, int32_t* ws_rows, double* ws_vals, double* ws_counts, double* out, void* stream) {
    SFB_REQUIRE(n_rows >= 0 && n_strains >= 0, "negative size");
    if (n_strains == 0) return SFB_OK;
    SFB_REQUIRE(out && ws_counts, "null output");
    SFB_REQUIRE(n_rows == 0 || (row_off && row_strain && row_count && mam && mam_nz && ratio_cov && ws_rows && ws_vals),
                "null array");
    cudaStream_t st = (cudaStream_t)stream;
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

}  // extern "C"
