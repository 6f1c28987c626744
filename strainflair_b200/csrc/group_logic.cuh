// Decision logic of the two passes, shared by the global-memory kernels (group.cu) and the tile kernels (tile.cu).
//
// Written against two small interfaces:
//   V    a view of a read group's match tuples ("colored_match", json2csv.py:851-866): nt(k), pid(k,i), code(k,i),
//        aloc(k,i) for mate k, tuple i, in the reference's order (LocalView / GlobalView in common.cuh, TileView);
//   Ops  where the effects go: assign_unique / hist / first_cluster / book (pass 1), uniq / apply (pass 2) -- global
//        atomics for the global kernels, shared-memory bins of the resident tile for the tile kernels.
#pragma once
#include "common.cuh"

namespace sfb {

// Strain reduction (json2csv.py:964-986): size of each mate's reduced list (a tuple counts once per
// common strain its path carries) and whether any common strain exists.
struct StrainCut {
    bool any;
    int n_new[2];
    u64 common0;   // the common-strain word when mask_words == 1
};
template <class V>
__device__ __forceinline__ StrainCut strain_cut(const sfb_graph& g, const V& v) {
    StrainCut c;
    c.any = false;
    c.n_new[0] = c.n_new[1] = 0;
    c.common0 = 0;
    for (int word = 0; word < (int)g.mask_words; ++word) {
        const u64 common = common_strains(g, v, word);
        if (word == 0) c.common0 = common;
        if (!common) continue;
        c.any = true;
        for (int k = 0; k < 2; ++k)
            for (int i = 0; i < v.nt(k); ++i)
                c.n_new[k] += __popcll(g.path_smask[(size_t)v.pid(k, i) * g.mask_words + word] & common);
    }
    return c;
}
template <class V>
__device__ __forceinline__ int times_of(const sfb_graph& g, const V& v, const StrainCut& c, int p) {
    if (g.mask_words == 1) return __popcll(g.path_smask[p] & c.common0);
    return strain_times(g, v, p);
}

// both mates have colored paths (json2csv.py:918-1000)
template <class V, class Ops>
__device__ __forceinline__ void classify_pair(const Dev& d, Ops& ops, BlockCounters& sc, const V& v, const Mate* mt,
                                              int* code, bool& defer) {
    int n_inter = 0, p_inter = -1;
    for (int i = 0; i < v.nt(0); ++i)
        if (first_of_pid(v, 0, i) && count_pid(v, 1, v.pid(0, i)) > 0) {
            if (n_inter++ == 0) p_inter = v.pid(0, i);
        }
    if (n_inter > 1) {
        SFB_COUNT(sc, SFB_C_SECOND_PASS, 2);
        defer = true;
        code[0] = code[1] = SFB_DEFER;
        return;
    }
    if (n_inter == 1) {
        for (int k = 0; k < 2; ++k) {
            if (count_pid(v, k, p_inter) > 1) {
                SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
                code[k] = SFB_MULTI_ALIGN;
                continue;
            }
            for (int i = 0; i < v.nt(k); ++i)
                if (v.pid(k, i) == p_inter) {
                    ops.assign_unique(d, sc, mt[k].a0 + v.aloc(k, i), v.code(k, i), p_inter, mt[k].seq_len);
                    break;
                }
            SFB_COUNT(sc, SFB_C_ASSIGNED_U, 1);
            SFB_COUNT(sc, SFB_C_SAME_CP, 1);
            if (v.nt(k) > 1) SFB_COUNT(sc, SFB_C_AMBIG_CP_RESOLVED, 1);
            code[k] = SFB_ASSIGN_SAME;
        }
        return;
    }
    // no common path: intersect strains (json2csv.py:961-1000)
    const StrainCut cut = strain_cut(d.g, v);
    if (!cut.any) {
        SFB_COUNT(sc, SFB_C_INCONSIST, 2);
        code[0] = code[1] = SFB_INCONSIST;
        return;
    }
    for (int k = 0; k < 2; ++k) {
        const int nn = cut.n_new[k];
        if (nn < v.nt(k)) SFB_COUNT(sc, nn == 1 ? SFB_C_AMBIG_RESOLVED : SFB_C_AMBIG_REDUCED, 1);
        if (nn > 1) {
            SFB_COUNT(sc, SFB_C_SECOND_PASS, 1);
            defer = true;
            code[k] = SFB_DEFER;
            continue;
        }
        for (int i = 0; i < v.nt(k); ++i) {      // exactly one tuple survives, once
            const int p = v.pid(k, i);
            if (times_of(d.g, v, cut, p) == 0) continue;
            if (count_pid(v, k, p) == 1) {
                ops.assign_unique(d, sc, mt[k].a0 + v.aloc(k, i), v.code(k, i), p, mt[k].seq_len);
                SFB_COUNT(sc, SFB_C_ASSIGNED_U, 1);
                SFB_COUNT(sc, SFB_C_DIFF_CP, 1);
                code[k] = SFB_ASSIGN_DIFF;
            } else {
                SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
                code[k] = SFB_MULTI_ALIGN;
            }
            break;
        }
    }
}


// both mates aligned, neither has a colored path (json2csv.py:868-891)
template <class Ops>
__device__ __forceinline__ void classify_both_unassigned(const Dev& d, Ops& ops, BlockCounters& sc, const Mate* mt, int* code) {
    SFB_COUNT(sc, SFB_C_UNASSIGNED, 2);
    int n_common = 0, common = 0;
    bool bad = false;
    for (int i = 0; i < mt[0].na; ++i) {
        const int ci = ops.first_cluster(d, mt[0].a0 + i);
        bad |= ci == -2;
        bool seen = false;
        for (int j = 0; j < i; ++j) seen |= ops.first_cluster(d, mt[0].a0 + j) == ci;
        if (seen) continue;
        bool other = false;
        for (int j = 0; j < mt[1].na; ++j) other |= ops.first_cluster(d, mt[1].a0 + j) == ci;
        if (other) { ++n_common; common = ci; }
    }
    for (int j = 0; j < mt[1].na; ++j) bad |= ops.first_cluster(d, mt[1].a0 + j) == -2;
    code[0] = code[1] = SFB_UNASSIGNED;
    if (bad) { SFB_COUNT(sc, SFB_C_NO_PATH, 1); return; }
    if (n_common > 1) return;
    for (int k = 0; k < 2; ++k) {
        int len = mt[k].na, cl = ops.first_cluster(d, mt[k].a0);
        if (n_common == 1) {
            len = 0;
            for (int j = 0; j < mt[k].na; ++j) len += ops.first_cluster(d, mt[k].a0 + j) == common;
            cl = common;
        }
        if (len == 1) { ops.book(d, sc, mt[k], cl); code[k] = SFB_UNASSIGNED_BOOK; }
    }
}

enum { kNone = 0, kInter = 1, kStrain = 2, kSingle = 3 };

struct Plan {
    int mode[2];
    int napp[2];
    int n_inter;
    long long total_inter;
    long long total_strain[2];
    int nn[2];
    StrainCut cut;
};

__device__ __forceinline__ double ratio_of(long long u, long long total, int n) {
    return total == 0 ? 1.0 / (double)n : (double)u / (double)total;        // json2csv.py:1151-1154
}

// json2csv.py:1123-1217: decide per mate; counters and grp_code2 here, the additions in p2_emit
template <class V, class Ops>
__device__ __forceinline__ void p2_decide_pair(const Dev& d, Ops& ops, BlockCounters& sc, const V& v, Plan& pl, int* code) {
    pl.n_inter = 0;
    pl.total_inter = 0;
    for (int i = 0; i < v.nt(0); ++i)
        if (first_of_pid(v, 0, i) && count_pid(v, 1, v.pid(0, i)) > 0) {
            ++pl.n_inter;
            pl.total_inter += ops.uniq(v.pid(0, i));
        }
    if (pl.n_inter >= 1) {
        for (int k = 0; k < 2; ++k) {
            bool simple = true;      // every common path has exactly one tuple in this mate
            for (int i = 0; i < v.nt(k) && simple; ++i) {
                const int p = v.pid(k, i);
                if (count_pid(v, 1 - k, p) > 0 && count_pid(v, k, p) != 1) simple = false;
            }
            if (!simple) {
                SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
                code[k] = SFB_P2_MULTI_ALIGN;
                continue;
            }
            SFB_COUNT(sc, SFB_C_ASSIGNED_M, 1);
            code[k] = SFB_P2_ASSIGNED;
            pl.mode[k] = kInter;
            pl.napp[k] = pl.n_inter;
        }
        return;
    }
    pl.cut = strain_cut(d.g, v);
    if (!pl.cut.any) {
        SFB_COUNT(sc, SFB_C_INCONSIST_M, (v.nt(0) == 1 || v.nt(1) == 1) ? 1 : 2);
        code[0] = code[1] = SFB_P2_INCONSIST;
        return;
    }
    for (int k = 0; k < 2; ++k) {
        pl.nn[k] = pl.cut.n_new[k];
        if (pl.nn[k] <= 1) continue;          // handled in pass 1 (json2csv.py:1185)
        bool simple = true;
        long long tot = 0;
        int napp = 0;
        for (int i = 0; i < v.nt(k); ++i) {
            const int p = v.pid(k, i);
            const int times = times_of(d.g, v, pl.cut, p);
            if (times == 0) continue;
            if (count_pid(v, k, p) != 1) { simple = false; break; }
            tot += ops.uniq(p) * times;
            ++napp;
        }
        if (!simple) {
            SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
            code[k] = SFB_P2_MULTI_ALIGN;
            continue;
        }
        SFB_COUNT(sc, SFB_C_ASSIGNED_M, 1);
        code[k] = SFB_P2_ASSIGNED;
        pl.mode[k] = kStrain;
        pl.napp[k] = napp;
        pl.total_strain[k] = tot;
    }
}

template <class V, class Ops>
__device__ __forceinline__ void p2_emit(const Dev& d, Ops& ops, BlockCounters& sc, const V& v, const Mate* mt, const Plan& pl) {
    for (int k = 0; k < 2; ++k) {
        if (pl.mode[k] == kInter) {
            for (int i = 0; i < v.nt(k); ++i) {
                const int p = v.pid(k, i);
                if (count_pid(v, 1 - k, p) > 0)
                    ops.apply(d, mt[k].a0 + v.aloc(k, i), v.code(k, i), p, mt[k].seq_len,
                             ratio_of(ops.uniq(p), pl.total_inter, pl.n_inter), 1);
            }
        } else if (pl.mode[k] == kStrain) {
            for (int i = 0; i < v.nt(k); ++i) {
                const int p = v.pid(k, i);
                const int times = times_of(d.g, v, pl.cut, p);
                if (times)
                    ops.apply(d, mt[k].a0 + v.aloc(k, i), v.code(k, i), p, mt[k].seq_len,
                             ratio_of(ops.uniq(p), pl.total_strain[k], pl.nn[k]), times);
            }
        } else if (pl.mode[k] == kSingle) {
            // every alignment separately (json2csv.py:1230-1268)
            int i = 0;
            while (i < v.nt(k)) {
                const int j = v.aloc(k, i);
                int e = i;
                long long total = 0;
                for (; e < v.nt(k) && v.aloc(k, e) == j; ++e) {
                    total += ops.uniq(v.pid(k, e));
                    ops.hist(d, sc, v.pid(k, e), mt[k].a0 + j);
                }
                for (int t = i; t < e; ++t)
                    ops.apply(d, mt[k].a0 + j, v.code(k, t), v.pid(k, t), mt[k].seq_len,
                             ratio_of(ops.uniq(v.pid(k, t)), total, e - i), 1);
                i = e;
            }
        }
    }
}


// single-end bookkeeping shared by both resolve kernels (json2csv.py:1270-1274)
__device__ __forceinline__ int p2_single_outcome(BlockCounters& sc, int n_tuples) {
    if (n_tuples > 0) {
        SFB_COUNT(sc, SFB_C_ASSIGNED_M, 1);
        SFB_COUNT(sc, SFB_C_SE_M, 1);
        return SFB_P2_SE_ASSIGNED;
    }
    SFB_COUNT(sc, SFB_C_UNASSIGNED, 1);
    return SFB_P2_UNASSIGNED;
}

}  // namespace sfb
