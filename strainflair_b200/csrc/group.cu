// Per-read-group logic: k_classify (pass-1 decision tree, json2csv.py:828-1053) and k_pass2_resolve
// (pass-2 candidate selection and ratios, json2csv.py:1085-1274).
//
// One thread per read group.  A group's match tuples are first gathered into thread-local arrays
// (LocalView: one coalesced 8-byte word per alignment, one contiguous list per multi-matched
// alignment); the set logic then runs on registers/L1 instead of chasing global memory.  Groups with
// more than kCap tuples per mate take the same code through GlobalView.
#include "common.cuh"
#include "group_fast.cuh"
#include "group_logic.cuh"
#include "group_ops.cuh"

namespace sfb {

// Effects of the decision logic (group_logic.cuh) for the global-memory kernels: atomics on the accumulators in HBM,
// pass-2 work items into app_buf (the node loops run in k_pass1_scatter / k_pass2_scatter).
struct GlobalOps {
    const long long* U;
    App* out;
    __device__ __forceinline__ void assign_unique(const Dev& d, BlockCounters& sc, i64 a, uint32_t code, int pid, int seq_len) {
        sfb::assign_unique(d, sc, a, code, pid, seq_len);
    }
    __device__ __forceinline__ void hist(const Dev& d, BlockCounters& sc, int pid, i64 a) { hist_add(d, sc, pid, a); }
    __device__ __forceinline__ int first_cluster(const Dev& d, i64 a) { return sfb::first_cluster(d, a); }
    __device__ __forceinline__ void book(const Dev& d, BlockCounters& sc, const Mate& mt, int cluster) { sfb::book(d, sc, mt, cluster); }
    __device__ __forceinline__ long long uniq(int pid) const { return U[pid]; }
    __device__ __forceinline__ void apply(const Dev& d, i64 a, uint32_t code, int pid, int seq_len, double ratio, int times) {
        emit_app(d, out++, a, code, pid, seq_len, ratio, times);
    }
};

}  // namespace sfb
#include "group_sets.cuh"
namespace sfb {

// the work item of a share found on path sets: it names its path, k_pass2_scatter finds the slot (group_sets.cuh)
struct AppEmit {
    const Dev& d;
    App* out;
    __device__ __forceinline__ void operator()(i64 a, bool rev, int pid, int seq_len, double ratio, int times) {
        emit_app(d, out++, a, kAppPath | (rev ? SFB_MATCH_REV : 0u) | (uint32_t)pid, pid, seq_len, ratio, times);
    }
};

// =============================================================================================
// classify
// =============================================================================================
// ---- register fast path of classify_pair (see group_fast.cuh) ---------------------------------
__device__ __forceinline__ void strain_mate_fast(const Dev& d, BlockCounters& sc, const Mate& mt, const RegMate& R,
                                                 const u64* sw, u64 common, unsigned dup, int& code, bool& defer) {
    int nn = 0;
    unsigned sel = 0;
#pragma unroll
    for (int i = 0; i < kR; ++i) {
        const int t = __popcll(sw[i] & common);
        nn += t;
        sel |= (unsigned)(t > 0) << i;
    }
    if (nn < R.n) SFB_COUNT(sc, nn == 1 ? SFB_C_AMBIG_RESOLVED : SFB_C_AMBIG_REDUCED, 1);
    if (nn > 1) {
        SFB_COUNT(sc, SFB_C_SECOND_PASS, 1);
        defer = true;
        code = SFB_DEFER;
        return;
    }
    const int i = __ffs(sel) - 1;            // the one tuple that survives, once
    if ((dup >> i) & 1u) {
        SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
        code = SFB_MULTI_ALIGN;
        return;
    }
    assign_unique(d, sc, mt.a0 + (i >= R.c0), reg_pick(R.code, i), reg_pick(R.pid, i), mt.seq_len);
    SFB_COUNT(sc, SFB_C_ASSIGNED_U, 1);
    SFB_COUNT(sc, SFB_C_DIFF_CP, 1);
    code = SFB_ASSIGN_DIFF;
}

__device__ __forceinline__ void classify_pair_fast(const Dev& d, BlockCounters& sc, const Mate* mt, const RegMate& A,
                                                   const RegMate& B, int* code, bool& defer) {
    const PairMasks pm = pair_masks(A, B);
    const unsigned inter0 = pm.in_other[0] & pm.first[0];
    const int n_inter = __popc(inter0);
    if (n_inter > 1) {
        SFB_COUNT(sc, SFB_C_SECOND_PASS, 2);
        defer = true;
        code[0] = code[1] = SFB_DEFER;
        return;
    }
    if (n_inter == 1) {
        const int i0 = __ffs(inter0) - 1;
        const int p = reg_pick(A.pid, i0);
        unsigned jm = 0;
#pragma unroll
        for (int j = 0; j < kR; ++j) jm |= (unsigned)(B.pid[j] == p) << j;
        const int j0 = __ffs(jm) - 1;
        const bool multi[2] = {((pm.dup[0] >> i0) & 1u) != 0, __popc(jm) > 1};
        const int idx[2] = {i0, j0};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const RegMate& R = k ? B : A;
            if (multi[k]) {
                SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
                code[k] = SFB_MULTI_ALIGN;
                continue;
            }
            assign_unique(d, sc, mt[k].a0 + (idx[k] >= R.c0), reg_pick(R.code, idx[k]), p, mt[k].seq_len);
            SFB_COUNT(sc, SFB_C_ASSIGNED_U, 1);
            SFB_COUNT(sc, SFB_C_SAME_CP, 1);
            if (R.n > 1) SFB_COUNT(sc, SFB_C_AMBIG_CP_RESOLVED, 1);
            code[k] = SFB_ASSIGN_SAME;
        }
        return;
    }
    u64 s0[kR], s1[kR];
    strain_words(d.g, A, s0);
    strain_words(d.g, B, s1);
    const u64 common = or_all(s0) & or_all(s1);
    if (!common) {
        SFB_COUNT(sc, SFB_C_INCONSIST, 2);
        code[0] = code[1] = SFB_INCONSIST;
        return;
    }
    strain_mate_fast(d, sc, mt[0], A, s0, common, pm.dup[0], code[0], defer);
    strain_mate_fast(d, sc, mt[1], B, s1, common, pm.dup[1], code[1], defer);
}

// single-end style (json2csv.py:1003-1050)
__device__ __forceinline__ int classify_single(const Dev& d, BlockCounters& sc, const Mate& mt, bool& defer) {
    SFB_COUNT(sc, SFB_C_SINGLES, 1);
    if (mt.na == 0) { SFB_COUNT(sc, SFB_C_FILTERED, 1); return SFB_FILTERED; }
    if (mt.na > 1) { SFB_COUNT(sc, SFB_C_SECOND_PASS, 1); defer = true; return SFB_DEFER; }
    const int n = match_count_at(d.w, mt.a0);
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

// one mate without colored path (json2csv.py:895-916); returns the mate that continues single-end style
__device__ __forceinline__ int classify_one_without_path(const Dev& d, BlockCounters& sc, const Mate* mt, int gone, int* code) {
    SFB_COUNT(sc, SFB_C_UNASSIGNED, 1);
    SFB_COUNT(sc, SFB_C_ONE_NOCP, 1);
    code[gone] = SFB_UNASSIGNED;
    bool bad = false;
    for (int j = 0; j < mt[gone].na; ++j) bad |= first_cluster(d, mt[gone].a0 + j) == -2;
    if (bad) {
        SFB_COUNT(sc, SFB_C_NO_PATH, 1);
    } else if (mt[gone].na == 1) {
        book(d, sc, mt[gone], first_cluster(d, mt[gone].a0));
        code[gone] = SFB_UNASSIGNED_BOOK;
    }
    return 1 - gone;
}

// Generic handling of a pair whose mates both have alignments (json2csv.py:847-1000); `single` is set
// when one mate has no colored path and the other continues single-end style.
__device__ __forceinline__ void classify_aligned_pair_generic(const Dev& d, BlockCounters& sc, const Mate* mt, int* code,
                                                              bool& defer, int& single) {
    LocalView lv;
    GlobalView gv;
    const bool fits = lv.load(d.w, mt);
    if (!fits) gv.init(d.w, mt);
    const int nt0 = fits ? lv.nt(0) : gv.nt(0), nt1 = fits ? lv.nt(1) : gv.nt(1);
    if (nt0 == 0 && nt1 == 0) {
        GlobalOps ops{nullptr, nullptr};
        classify_both_unassigned(d, ops, sc, mt, code);
    } else if (nt0 == 0 || nt1 == 0) {
        single = classify_one_without_path(d, sc, mt, nt0 == 0 ? 0 : 1, code);
    } else if (fits) {
        GlobalOps ops{nullptr, nullptr};
        classify_pair(d, ops, sc, lv, mt, code, defer);
    } else {
        GlobalOps ops{nullptr, nullptr};
        classify_pair(d, ops, sc, gv, mt, code, defer);
    }
}

#ifndef SFB_CLS_PREFETCH
#define SFB_CLS_PREFETCH 1
#endif
#ifndef SFB_CLS_MINBLK
#define SFB_CLS_MINBLK 8      /* 80 registers; measured, see profiles/README.md */
#endif
#ifndef SFB_P2_MINBLK
#define SFB_P2_MINBLK 6       /* 64 registers */
#endif
// Fast kernel: everything except aligned pairs that do not fit the register path, which go to w.slow.  SETS: the matches
// are the path sets of sfb_match_steps (group_sets.cuh), else tuples (group_fast.cuh); one kernel per form, so that
// neither carries the other's registers.
template <bool SETS>
__global__ void __launch_bounds__(128, SFB_CLS_MINBLK) k_classify(Dev d) {
    __shared__ BlockCounters sc;
    block_counters_zero(sc);
    __syncthreads();
    const i64 n_round = d.w.grp_begin + ((d.r.n_groups - d.w.grp_begin + 127) & ~(i64)127);
    const i64 stride = (i64)gridDim.x * blockDim.x;
    // two passes ahead: the first alignment of the group this thread takes then; one pass ahead: its match words, path
    // sets and walk offsets on their way to the L1 -- the chain header -> match word -> sets -> walk offset is four of the
    // seven dependent gathers of a group
    i64 a_next = -1;
    {
        const i64 g1 = d.w.grp_begin + (i64)blockIdx.x * blockDim.x + threadIdx.x + stride;
        if (SFB_CLS_PREFETCH && g1 < d.r.n_groups) a_next = d.r.grp_aln_off[2 * g1];
    }
    for (i64 g = d.w.grp_begin + (i64)blockIdx.x * blockDim.x + threadIdx.x; g < n_round; g += stride) {
        if (SFB_CLS_PREFETCH) {
            if (a_next >= 0 && a_next < d.r.n_alns) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const uint2*>(d.w.aln_match) + a_next));
                if (SETS) asm volatile("prefetch.global.L1 [%0];" ::"l"(d.w.aln_sets + 2 * a_next));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(d.r.aln_walk_off + a_next));
            }
            const i64 g2 = g + 2 * stride;
            a_next = g2 < d.r.n_groups ? d.r.grp_aln_off[2 * g2] : -1;
            if (g2 < d.r.n_groups) asm volatile("prefetch.global.L1 [%0];" ::"l"(d.r.mate_seq_len + 2 * g2));
        }
        bool defer = false, slow = false;
        Mate mt[2];
        SetMate sm[2];
        int want[2] = {-1, -1};                  // SETS: the path bit each mate is assigned to (sets_assign_pair below)
        mt[0].seq_len = mt[1].seq_len = 0;
        sm[0].p0 = sm[1].p0 = -1;
        if (g < d.r.n_groups) {
            const int nm = d.r.grp_nmates[g];
            mate_load(d.r, g, 0, mt[0]);
            mate_load(d.r, g, 1, mt[1]);
            for (int k = 0; k < 2; ++k)
                for (int j = 0; j < mt[k].na; ++j) d.w.aln_assign[mt[k].a0 + j] = -1;
            int code[2] = {SFB_ABSENT, SFB_ABSENT};
            SFB_COUNT(sc, SFB_C_NB_READS, nm);
            int single = nm == 1 ? 0 : -1;       // mate handled single-end style, -1 = none
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
                } else if (SETS) {
                    // path sets of the step index: any number of candidate paths in registers
                    if (d.g.mask_words == 1 && sets_load(d.w, mt[0], sm[0]) && sets_load(d.w, mt[1], sm[1])) {
                        if (sm[0].nt == 0 && sm[1].nt == 0) {
                            GlobalOps ops{nullptr, nullptr};
                            classify_both_unassigned(d, ops, sc, mt, code);
                        } else if (sm[0].nt == 0) {
                            single = classify_one_without_path(d, sc, mt, 0, code);
                        } else if (sm[1].nt == 0) {
                            single = classify_one_without_path(d, sc, mt, 1, code);
                        } else {
                            sets_classify_pair(d, sc, mt, sm, code, defer, want);
                        }
                    } else {
                        slow = true;
                    }
                } else {
                    RegMate A, B;
                    const bool quick = d.g.mask_words == 1 && reg_load(d.w, mt[0], -2, A) && reg_load(d.w, mt[1], -32, B) &&
                                       A.n > 0 && B.n > 0;
                    if (quick) classify_pair_fast(d, sc, mt, A, B, code, defer);
                    else slow = true;
                }
            }
            if (!slow) {
                if (single >= 0) code[single] = classify_single(d, sc, mt[single], defer);
                d.w.grp_code[g] = (uint8_t)(code[0] | (code[1] << 4));
            }
        }
        // the unique assignments the tree noted, both mates side by side, the whole warp together
        if (SETS && __any_sync(0xffffffffu, want[0] >= 0 || want[1] >= 0)) sets_assign_pair(d, sc, mt, sm, want);
        // warp-aggregated appends: deferred groups (pass 2; those before p2_begin are found by sfb_tile_share through
        // their grp_code and only counted) and groups for the generic kernel
        const bool list = defer && g >= d.w.p2_begin;
        const u64 pos = warp_claim(d.w.cursor + 1, list ? 1u : 0u);
        if (list) d.w.deferred[pos] = (int32_t)g;
        const unsigned held = __ballot_sync(0xffffffffu, defer && !list);
        if (held && lane_id() == 0) atomicAdd(d.w.cursor + 7, (u64)__popc(held));
        const u64 spos = warp_claim(d.w.cursor + 3, slow ? 1u : 0u);
        if (slow) d.w.slow[spos] = (int32_t)g;
    }
    __syncthreads();
    block_counters_flush(sc, d.acc.counters);
}

// Generic kernel: one thread per group of the slow list (measured faster than a warp per group for pass 1, where
// most of the work is a handful of set tests; pass 2 is the other way round, see group_warp.cuh).
__global__ void __launch_bounds__(128) k_classify_slow(Dev d) {
    __shared__ BlockCounters sc;
    block_counters_zero(sc);
    __syncthreads();
    const i64 n_slow = (i64)d.w.cursor[3];
    const i64 n_round = (n_slow + 127) & ~(i64)127;
    for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < n_round; q += (i64)gridDim.x * blockDim.x) {
        bool defer = false;
        i64 g = 0;
        if (q < n_slow) {
            g = d.w.slow[q];
            Mate mt[2];
            mate_load(d.r, g, 0, mt[0]);
            mate_load(d.r, g, 1, mt[1]);
            int code[2] = {SFB_ABSENT, SFB_ABSENT};
            int single = -1;
            classify_aligned_pair_generic(d, sc, mt, code, defer, single);
            if (single >= 0) code[single] = classify_single(d, sc, mt[single], defer);
            d.w.grp_code[g] = (uint8_t)(code[0] | (code[1] << 4));
        }
        const bool list = defer && g >= d.w.p2_begin;
        const u64 pos = warp_claim(d.w.cursor + 1, list ? 1u : 0u);
        if (list) d.w.deferred[pos] = (int32_t)g;
        const unsigned held = __ballot_sync(0xffffffffu, defer && !list);
        if (held && lane_id() == 0) atomicAdd(d.w.cursor + 7, (u64)__popc(held));
    }
    __syncthreads();
    block_counters_flush(sc, d.acc.counters);
}

// =============================================================================================
// pass 2, resolve: which (alignment, path position) pairs receive which share
// =============================================================================================
// ---- register fast path of pass 2 (see group_fast.cuh) -----------------------------------------
struct FastPlan {
    int mode[2];
    unsigned sel[2];          // tuples that receive a share
    int n_inter;
    long long total_inter;
    long long total_strain[2];
    int nn[2];
    u64 common;
    long long total_single[2];   // per alignment of the single-end mate
};

__device__ __forceinline__ void p2_fast_decide_pair(const Dev& d, BlockCounters& sc, const RegMate& A, const RegMate& B,
                                                    const long long* U, FastPlan& pl, int* code) {
    const PairMasks pm = pair_masks(A, B);
    const unsigned inter0 = pm.in_other[0] & pm.first[0];
    pl.n_inter = __popc(inter0);
    if (pl.n_inter >= 1) {
        long long total = 0;
        for (unsigned rest = inter0; rest; rest &= rest - 1) total += U[reg_pick(A.pid, __ffs(rest) - 1)];
        pl.total_inter = total;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (pm.in_other[k] & pm.dup[k]) {       // a common path with several tuples in this mate
                SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
                code[k] = SFB_P2_MULTI_ALIGN;
                continue;
            }
            SFB_COUNT(sc, SFB_C_ASSIGNED_M, 1);
            code[k] = SFB_P2_ASSIGNED;
            pl.mode[k] = kInter;
            pl.sel[k] = pm.in_other[k];
        }
        return;
    }
    u64 s0[kR], s1[kR];
    strain_words(d.g, A, s0);
    strain_words(d.g, B, s1);
    pl.common = or_all(s0) & or_all(s1);
    if (!pl.common) {
        SFB_COUNT(sc, SFB_C_INCONSIST_M, (A.n == 1 || B.n == 1) ? 1 : 2);
        code[0] = code[1] = SFB_P2_INCONSIST;
        return;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const RegMate& R = k ? B : A;
        const u64* sw = k ? s1 : s0;
        int nn = 0;
        unsigned sel = 0;
        long long tot = 0;
#pragma unroll
        for (int i = 0; i < kR; ++i) {
            const int t = __popcll(sw[i] & pl.common);
            nn += t;
            sel |= (unsigned)(t > 0) << i;
        }
        pl.nn[k] = nn;
        if (nn <= 1) continue;                       // handled in pass 1 (json2csv.py:1185)
        if (sel & pm.dup[k]) {
            SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
            code[k] = SFB_P2_MULTI_ALIGN;
            continue;
        }
        for (unsigned rest = sel; rest; rest &= rest - 1) {
            const int i = __ffs(rest) - 1;
            tot += U[reg_pick(R.pid, i)] * __popcll(reg_pick(sw, i) & pl.common);
        }
        SFB_COUNT(sc, SFB_C_ASSIGNED_M, 1);
        code[k] = SFB_P2_ASSIGNED;
        pl.mode[k] = kStrain;
        pl.sel[k] = sel;
        pl.total_strain[k] = tot;
    }
}

__device__ __forceinline__ void p2_fast_single(const Dev& d, BlockCounters& sc, const RegMate& R, const long long* U,
                                               FastPlan& pl, int k) {
    long long t0 = 0, t1 = 0;
    for (int i = 0; i < R.n; ++i) {
        const long long u = U[reg_pick(R.pid, i)];
        if (i < R.c0) t0 += u; else t1 += u;
    }
    pl.total_single[0] = t0;
    pl.total_single[1] = t1;
    pl.mode[k] = kSingle;
    pl.sel[k] = (1u << R.n) - 1u;
}

__device__ __forceinline__ App* p2_fast_emit(const Dev& d, BlockCounters& sc, const Mate& mt, const RegMate& R, int mode,
                                             unsigned sel, const FastPlan& pl, int k, const long long* U, App* out) {
    for (unsigned rest = sel; rest; rest &= rest - 1) {
        const int i = __ffs(rest) - 1;
        const int pid = reg_pick(R.pid, i);
        const long long u = U[pid];
        const int second = i >= R.c0;
        double w;
        int times = 1;
        if (mode == kInter) {
            w = ratio_of(u, pl.total_inter, pl.n_inter);
        } else if (mode == kStrain) {
            times = __popcll(d.g.path_smask[pid] & pl.common);
            w = ratio_of(u, pl.total_strain[k], pl.nn[k]);
        } else {
            hist_add(d, sc, pid, mt.a0 + second);                                  // json2csv.py:1246-1252
            w = ratio_of(u, pl.total_single[second], second ? R.n - R.c0 : R.c0);
        }
        emit_app(d, out++, mt.a0 + second, reg_pick(R.code, i), pid, mt.seq_len, w, times);
    }
    return out;
}

// Fast kernel over the deferred groups; groups that do not fit the register path go to w.slow (cursor[4]).
template <bool SETS>
__global__ void __launch_bounds__(128, SFB_P2_MINBLK) k_pass2_resolve(Dev d, const long long* __restrict__ U) {
    __shared__ BlockCounters sc;
    block_counters_zero(sc);
    __syncthreads();
    const i64 n_def = (i64)d.w.cursor[1];
    const i64 n_round = (n_def + 127) & ~(i64)127;
    App* apps = reinterpret_cast<App*>(d.w.app_buf);
    for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < n_round; q += (i64)gridDim.x * blockDim.x) {
        const bool live = q < n_def;
        RegMate A, B;
        FastPlan fp;
        fp.mode[0] = fp.mode[1] = kNone;
        fp.sel[0] = fp.sel[1] = 0;
        SetMate sm[2];
        SetPlan sp;
        sp.mode[0] = sp.mode[1] = kNone;
        Mate mt[2];
        bool slow = false, sets = false;
        i64 g = 0;
        unsigned need = 0;
        if (live) {
            g = d.w.deferred[q];
            const int nm = d.r.grp_nmates[g];
            mate_load(d.r, g, 0, mt[0]);
            mate_load(d.r, g, 1, mt[1]);
            sets = SETS && d.g.mask_words == 1 && sets_load(d.w, mt[0], sm[0]) && sets_load(d.w, mt[1], sm[1]);
            const bool quick = !SETS && d.g.mask_words == 1 && reg_load(d.w, mt[0], -2, A) && reg_load(d.w, mt[1], -32, B);
            if (sets) {
                int code[2] = {SFB_P2_NONE, SFB_P2_NONE};
                int single = nm == 1 ? 0 : -1;
                if (nm == 2) {
                    if (mt[0].na == 0 || mt[1].na == 0) single = mt[0].na == 0 ? 1 : 0;          // json2csv.py:1088-1093
                    else if (sm[0].nt == 0 || sm[1].nt == 0) single = sm[0].nt == 0 ? 1 : 0;    // :1117-1122
                    else sets_p2_decide_pair(d, sc, sm, U, sp, code);
                }
                if (single == 0) {                    // (no run-time index into the register structs)
                    sp.mode[0] = kSingle;
                    code[0] = p2_single_outcome(sc, sm[0].nt);
                } else if (single == 1) {
                    sp.mode[1] = kSingle;
                    code[1] = p2_single_outcome(sc, sm[1].nt);
                }
                d.w.grp_code2[g] = (uint8_t)(code[0] | (code[1] << 4));
                need = (unsigned)(sets_p2_need(sm[0], sp.mode[0], sp.sel[0]) + sets_p2_need(sm[1], sp.mode[1], sp.sel[1]));
            } else if (!quick) {
                slow = true;
            } else {
                int code[2] = {SFB_P2_NONE, SFB_P2_NONE};
                int single = nm == 1 ? 0 : -1;
                if (nm == 2) {
                    if (mt[0].na == 0 || mt[1].na == 0) single = mt[0].na == 0 ? 1 : 0;      // json2csv.py:1088-1093
                    else if (A.n == 0 || B.n == 0) single = A.n == 0 ? 1 : 0;                // :1117-1122
                    else p2_fast_decide_pair(d, sc, A, B, U, fp, code);
                }
                if (single == 0) {
                    p2_fast_single(d, sc, A, U, fp, 0);
                    code[0] = p2_single_outcome(sc, A.n);
                } else if (single == 1) {
                    p2_fast_single(d, sc, B, U, fp, 1);
                    code[1] = p2_single_outcome(sc, B.n);
                }
                d.w.grp_code2[g] = (uint8_t)(code[0] | (code[1] << 4));
                need = (unsigned)(__popc(fp.sel[0]) + __popc(fp.sel[1]));
            }
        }
        const u64 pos = warp_claim(d.w.cursor + 2, need);
        if (live && need) {
            SFB_COUNT(sc, SFB_C_ST_P2_APPLIED, (int)need);
            if (pos + need <= (u64)d.w.app_cap) {
                if (SETS) {
                    AppEmit emit{d, apps + pos};
                    sets_p2_emit(d, sc, mt[0], sm[0], sp, 0, U, emit);
                    sets_p2_emit(d, sc, mt[1], sm[1], sp, 1, U, emit);
                } else {
                    App* out = p2_fast_emit(d, sc, mt[0], A, fp.mode[0], fp.sel[0], fp, 0, U, apps + pos);
                    p2_fast_emit(d, sc, mt[1], B, fp.mode[1], fp.sel[1], fp, 1, U, out);
                }
            } else {
                SFB_COUNT(sc, SFB_C_MATCH_OVERFLOW, (int)need);
            }
        }
        const u64 spos = warp_claim(d.w.cursor + 4, slow ? 1u : 0u);
        if (slow) d.w.slow[spos] = (int32_t)g;
    }
    __syncthreads();
    block_counters_flush(sc, d.acc.counters);
}

#include "group_warp.cuh"

int launch_expand_sets(const sfb_graph*, const sfb_reads*, const sfb_work*, sfb_accum*, int, cudaStream_t);

int launch_classify(const Dev& d, cudaStream_t st) {
    if (d.r.n_groups == 0) return SFB_OK;
    static OccCache c_fast, c_sets, c_slow;
    const bool sets = d.g.node_step && d.w.aln_sets;          // the matches came from sfb_match_steps
    if (d.r.n_groups > d.w.grp_begin) {
        if (sets) {
            const unsigned grid = persistent_grid(k_classify<true>, 128, blocks_for(d.r.n_groups - d.w.grp_begin, 128), c_sets);
            k_classify<true><<<grid, 128, 0, st>>>(d);
        } else {
            const unsigned grid = persistent_grid(k_classify<false>, 128, blocks_for(d.r.n_groups - d.w.grp_begin, 128), c_fast);
            k_classify<false><<<grid, 128, 0, st>>>(d);
        }
        SFB_TRY(check_launch("k_classify"));
    }
    if (sets) SFB_TRY(launch_expand_sets(&d.g, &d.r, &d.w, const_cast<sfb_accum*>(&d.acc), 3, st));   // tuple lists for the generic kernel
    const unsigned grid2 = persistent_grid(k_classify_slow, 128, blocks_for(d.r.n_groups, 128), c_slow);
    k_classify_slow<<<grid2, 128, 0, st>>>(d);             // usually a few percent of the groups
    return check_launch("k_classify_slow");
}
int launch_pass2_resolve(const Dev& d, const long long* U, cudaStream_t st) {
    if (d.r.n_groups == 0) return SFB_OK;
    static OccCache c_fast, c_sets, c_slow;
    const bool sets = d.g.node_step && d.w.aln_sets;
    if (sets) {
        const unsigned grid = persistent_grid(k_pass2_resolve<true>, 128, blocks_for(d.r.n_groups, 128), c_sets);
        k_pass2_resolve<true><<<grid, 128, 0, st>>>(d, U);
    } else {
        const unsigned grid = persistent_grid(k_pass2_resolve<false>, 128, blocks_for(d.r.n_groups, 128), c_fast);
        k_pass2_resolve<false><<<grid, 128, 0, st>>>(d, U);
    }
    SFB_TRY(check_launch("k_pass2_resolve"));
    // (a deferred group that never went through the generic classify kernel -- a single mate with several alignments --
    // may still hold path sets when it takes the generic kernel here)
    if (sets) SFB_TRY(launch_expand_sets(&d.g, &d.r, &d.w, const_cast<sfb_accum*>(&d.acc), 4, st));
    const unsigned grid2 = persistent_grid(k_pass2_resolve_slow, 128, blocks_for(d.r.n_groups, 128), c_slow);
    k_pass2_resolve_slow<<<grid2, 128, 0, st>>>(d, U);
    return check_launch("k_pass2_resolve_slow");
}

}  // namespace sfb
