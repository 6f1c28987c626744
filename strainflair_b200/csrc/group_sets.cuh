// The per-group logic on path sets (SFB_MATCH_SETS, sfb_match_steps): the fast path of k_classify / k_pass2_resolve for
// read groups whose alignments all lie in simple tiles.
//
// A mate with at most two retained alignments is four 64-bit sets in registers -- forward and reverse matches of each
// alignment, bit = path id - p0 of the tile -- whatever the number of candidate paths (the tuple fast path of
// group_fast.cuh holds eight tuples per mate).  "colored_match" (json2csv.py:851-866) lists the tuples alignment by
// alignment, forward matches by ascending path, then reverse: tuple order only matters for picking "the first tuple on
// path p", which is the first of the four sets that holds p.  Set algebra of json2csv.py:922-1000 / :1123-1217:
//     paths of a mate            S = OR of its sets          paths with several tuples   dup
//     common paths               S0 & S1 (same tile)         list.count(p)               number of sets holding p
//     strains of a mate          OR of path_smask over S     strain-reduced list         count(p) * popcount(smask[p] & common)
#pragma once
#include "common.cuh"
#include "group_logic.cuh"
#include "group_ops.cuh"

#ifndef SFB_CLS_AGG
#define SFB_CLS_AGG 1
#endif

namespace sfb {

struct SetMate {
    int na, p0, nt;      // alignments, first path id of the tile (-1: no match at all), tuples
    u64 m[4];            // forward / reverse set of the first alignment, of the second
    u64 S, dup;
};
// false: the mate does not fit (more than two alignments, tuple lists, alignments in different tiles)
__device__ __forceinline__ bool sets_load(const sfb_work& w, const Mate& mt, SetMate& s) {
    s.na = mt.na;
    s.p0 = -1;
    s.m[0] = s.m[1] = s.m[2] = s.m[3] = 0;
    if (mt.na > 2) return false;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 2; ++j)
        if (j < mt.na) {
            const uint2 x = match_word(w, mt.a0 + j);
            if (x.x == SFB_MATCH_SETS) {
                ok &= s.p0 < 0 || s.p0 == (int)x.y;
                s.p0 = (int)x.y;
                const ulonglong2 v = reinterpret_cast<const ulonglong2*>(w.aln_sets)[mt.a0 + j];
                s.m[2 * j] = v.x;
                s.m[2 * j + 1] = v.y;
            } else {
                ok &= x.x == SFB_MATCH_NONE;
            }
        }
    s.S = s.m[0] | s.m[1] | s.m[2] | s.m[3];
    s.dup = (s.m[0] & s.m[1]) | ((s.m[0] | s.m[1]) & s.m[2]) | ((s.m[0] | s.m[1] | s.m[2]) & s.m[3]);
    s.nt = __popcll(s.m[0]) + __popcll(s.m[1]) + __popcll(s.m[2]) + __popcll(s.m[3]);
    return ok;
}
__device__ __forceinline__ int sets_count(const SetMate& s, int p) {          // list_path_id.count(pid)
    return (int)((s.m[0] >> p) & 1ull) + (int)((s.m[1] >> p) & 1ull) + (int)((s.m[2] >> p) & 1ull) + (int)((s.m[3] >> p) & 1ull);
}
// the node that carries the tuples of set q of the mate: the walk's first node (forward sets), its last (reverse sets)
__device__ __forceinline__ int sets_node(const sfb_reads& r, const Mate& mt, int q) {
    const i64 a = mt.a0 + (q >> 1);
    return (q & 1) ? r.walk_node[r.aln_walk_off[a + 1] - 1] : r.walk_node[r.aln_walk_off[a]];
}
// fill_abund_dist (json2csv.py:321-339) for the first tuple of each mate on its path bit want[k] (-1: nothing to assign).
// The decision tree only NOTES what it assigns; the two assignments of a pair then run here side by side, level by level --
// walk offset -> node that carries the tuple -> node_mask / occ_off -> occurrence (= slot) are four dependent gathers, and
// inside the branches of the tree they ran one mate after the other, in whatever lanes had taken that branch.
__device__ __forceinline__ void sets_assign_pair(const Dev& d, BlockCounters& sc, const Mate* mt, const SetMate* sm, const int* want) {
    i64 a[2], at[2];
    int q[2], node[2], oo[2], plen[2], nstr[2];
    u64 nm[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        q[k] = 3;
        a[k] = at[k] = 0;
        if (want[k] < 0) continue;
#pragma unroll
        for (int j = 2; j >= 0; --j)
            if ((sm[k].m[j] >> want[k]) & 1ull) q[k] = j;
        a[k] = mt[k].a0 + (q[k] >> 1);
        at[k] = d.r.aln_walk_off[a[k] + (q[k] & 1)] - (q[k] & 1);      // the walk's first record, its last for a reverse match
        plen[k] = d.g.path_seq_len[sm[k].p0 + want[k]];
        nstr[k] = d.g.path_nstrain[sm[k].p0 + want[k]];
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) node[k] = want[k] >= 0 ? d.r.walk_node[at[k]] : 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        nm[k] = 0;
        oo[k] = 0;
        if (want[k] >= 0) {
            nm[k] = d.g.node_mask[node[k]];
            oo[k] = d.g.occ_off[node[k]];
        }
    }
#if SFB_CLS_AGG
    // uniq_reads / uniq_norm (json2csv.py:326-327): reads in tile order put neighbouring lanes on the same few paths, so
    // the lanes of the warp that add the same term to the same path send ONE update -- the integer term times their
    // number (exact: same sums as one update per lane); called by the whole warp
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const bool on = want[k] >= 0;
        const int pid = on ? sm[k].p0 + want[k] : -1 - lane_id();
        const unsigned peers = __match_any_sync(0xffffffffu, ((u64)(uint32_t)pid << 32) | (uint32_t)mt[k].seq_len);
        if (on && lane_id() == __ffs((int)peers) - 1) {
            const u64 cnt = (u64)__popc(peers);
            atomicAdd((u64*)&d.acc.uniq_reads[pid], cnt);
            const double v = (double)mt[k].seq_len / (double)plen[k];
            double* dst = &d.acc.uniq_norm[pid];
            if (d.acc.det_stride == 0) {
                atomicAdd(dst, v * (double)cnt);
            } else {
                const Fx f = fx_split(v);
                const u64 low = (u64)f.x0 * cnt;
                atomicAdd(reinterpret_cast<u64*>(dst), ((((u64)f.x2 << 32) | f.x1) * cnt) + (low >> 32));
                if ((uint32_t)low) atomicAdd(reinterpret_cast<u64*>(dst + d.acc.det_stride), (u64)(uint32_t)low);
            }
        }
    }
#endif
    int slot[2];
#pragma unroll
    for (int k = 0; k < 2; ++k)
        slot[k] = want[k] >= 0 ? d.g.occ_pair[2 * (size_t)(oo[k] + __popcll(nm[k] & ((1ull << want[k]) - 1)))] : 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (want[k] < 0) continue;
        const int pid = sm[k].p0 + want[k];
        const uint32_t code = (uint32_t)slot[k] | ((q[k] & 1) ? SFB_MATCH_REV : 0u);
        d.w.aln_assign[a[k]] = (int32_t)code;
#if !SFB_CLS_AGG
        atomicAdd((u64*)&d.acc.uniq_reads[pid], 1ull);
        acc_add(&d.acc.uniq_norm[pid], d.acc.det_stride, (double)mt[k].seq_len / (double)plen[k]);
#endif
        if (nstr[k] == 1) hist_add(d, sc, pid, a[k]);      // (aggregating these too -- ten __match_any_sync per pass of the warp -- is slower: 2.49 ms)
        if (a[k] < d.w.aln_begin) {
            // a read group of the tile range that sfb_tile_pass1 handed over: its node loop runs here (see assign_unique)
            const i64 w0 = d.r.aln_walk_off[a[k]];
            const int L = (int)(d.r.aln_walk_off[a[k] + 1] - w0);
            double* dst = d.acc.uniq_abund + (code & SFB_SLOT_MASK);
            for (int i = 0; i < L; ++i) acc_add(dst + i, d.acc.det_stride, record_cov(d.r, w0 + i));
            SFB_COUNT(sc, SFB_C_ST_P1_RECORDS, L);
        }
    }
}
// strains of the mate's candidate paths (mask_words == 1)
__device__ __forceinline__ u64 sets_strains(const sfb_graph& g, const SetMate& s) {
    u64 u = 0;
    for (u64 rest = s.S; rest; rest &= rest - 1) u |= g.path_smask[s.p0 + __ffsll((long long)rest) - 1];
    return u;
}
// size of the strain-reduced list: a tuple counts once per common strain its path carries (json2csv.py:964-986, Q4)
__device__ __forceinline__ int sets_reduced(const sfb_graph& g, const SetMate& s, u64 common) {
    int nn = 0;
    for (u64 rest = s.S; rest; rest &= rest - 1) {
        const int p = __ffsll((long long)rest) - 1;
        nn += sets_count(s, p) * __popcll(g.path_smask[s.p0 + p] & common);
    }
    return nn;
}

// both mates have colored paths (json2csv.py:918-1000); same decisions as classify_pair (group_logic.cuh)
// (`want`: the path bit each mate is assigned to, for sets_assign_pair; -1 = none)
__device__ __forceinline__ void sets_classify_pair(const Dev& d, BlockCounters& sc, const Mate* mt, const SetMate* sm, int* code,
                                                   bool& defer, int* want) {
    const u64 inter = sm[0].p0 == sm[1].p0 ? sm[0].S & sm[1].S : 0ull;
    const int n_inter = __popcll(inter);
    if (n_inter > 1) {
        SFB_COUNT(sc, SFB_C_SECOND_PASS, 2);
        defer = true;
        code[0] = code[1] = SFB_DEFER;
        return;
    }
    if (n_inter == 1) {
        const int p = __ffsll((long long)inter) - 1;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if ((sm[k].dup >> p) & 1ull) {
                SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
                code[k] = SFB_MULTI_ALIGN;
                continue;
            }
            want[k] = p;
            SFB_COUNT(sc, SFB_C_ASSIGNED_U, 1);
            SFB_COUNT(sc, SFB_C_SAME_CP, 1);
            if (sm[k].nt > 1) SFB_COUNT(sc, SFB_C_AMBIG_CP_RESOLVED, 1);
            code[k] = SFB_ASSIGN_SAME;
        }
        return;
    }
    // no common path: intersect strains (json2csv.py:961-1000)
    const u64 common = sets_strains(d.g, sm[0]) & sets_strains(d.g, sm[1]);
    if (!common) {
        SFB_COUNT(sc, SFB_C_INCONSIST, 2);
        code[0] = code[1] = SFB_INCONSIST;
        return;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int nn = sets_reduced(d.g, sm[k], common);
        if (nn < sm[k].nt) SFB_COUNT(sc, nn == 1 ? SFB_C_AMBIG_RESOLVED : SFB_C_AMBIG_REDUCED, 1);
        if (nn > 1) {
            SFB_COUNT(sc, SFB_C_SECOND_PASS, 1);
            defer = true;
            code[k] = SFB_DEFER;
            continue;
        }
        for (u64 rest = sm[k].S; rest; rest &= rest - 1) {       // exactly one tuple survives, once
            const int p = __ffsll((long long)rest) - 1;
            if ((d.g.path_smask[sm[k].p0 + p] & common) == 0) continue;
            if (sets_count(sm[k], p) == 1) {
                want[k] = p;
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

// single-end style (json2csv.py:1003-1050); same decisions as classify_single (group.cu)
__device__ __forceinline__ int sets_classify_single(const Dev& d, BlockCounters& sc, const Mate& mt, const SetMate& s, bool& defer) {
    SFB_COUNT(sc, SFB_C_SINGLES, 1);
    if (mt.na == 0) { SFB_COUNT(sc, SFB_C_FILTERED, 1); return SFB_FILTERED; }
    if (mt.na > 1 || s.nt > 1) { SFB_COUNT(sc, SFB_C_SECOND_PASS, 1); defer = true; return SFB_DEFER; }
    if (s.nt == 0) {
        SFB_COUNT(sc, SFB_C_UNASSIGNED, 1);
        const int c = first_cluster(d, mt.a0);
        if (c == -2) SFB_COUNT(sc, SFB_C_NO_PATH, 1); else book(d, sc, mt, c);
        return SFB_UNASSIGNED_BOOK;
    }
    SFB_COUNT(sc, SFB_C_ASSIGNED_U, 1);      // counted, never accumulated (json2csv.py:1047-1053, Q1)
    SFB_COUNT(sc, SFB_C_SE, 1);
    return SFB_SE_COUNTED;
}

// ---- pass 2 (json2csv.py:1085-1274) -------------------------------------------------------------------------------
struct SetPlan {
    int mode[2];           // kNone / kInter / kStrain / kSingle
    u64 sel[2];            // paths that receive a share
    long long total[2];    // kInter / kStrain: sum of the unique counts over the candidate list
    int n[2];              // ... and its length
    u64 common;
};
// json2csv.py:1123-1217; same decisions as p2_decide_pair (group_logic.cuh)
__device__ __forceinline__ void sets_p2_decide_pair(const Dev& d, BlockCounters& sc, const SetMate* sm, const long long* U,
                                                    SetPlan& pl, int* code) {
    const u64 inter = sm[0].p0 == sm[1].p0 ? sm[0].S & sm[1].S : 0ull;
    const int n_inter = __popcll(inter);
    if (n_inter >= 1) {
        long long total = 0;
        for (u64 rest = inter; rest; rest &= rest - 1) total += U[sm[0].p0 + __ffsll((long long)rest) - 1];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (sm[k].dup & inter) {             // a common path with several tuples in this mate
                SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
                code[k] = SFB_P2_MULTI_ALIGN;
                continue;
            }
            SFB_COUNT(sc, SFB_C_ASSIGNED_M, 1);
            code[k] = SFB_P2_ASSIGNED;
            pl.mode[k] = kInter;
            pl.sel[k] = inter;
            pl.total[k] = total;
            pl.n[k] = n_inter;
        }
        return;
    }
    pl.common = sets_strains(d.g, sm[0]) & sets_strains(d.g, sm[1]);
    if (!pl.common) {
        SFB_COUNT(sc, SFB_C_INCONSIST_M, (sm[0].nt == 1 || sm[1].nt == 1) ? 1 : 2);
        code[0] = code[1] = SFB_P2_INCONSIST;
        return;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int nn = sets_reduced(d.g, sm[k], pl.common);
        if (nn <= 1) continue;                   // handled in pass 1 (json2csv.py:1185)
        bool simple = true;
        long long tot = 0;
        u64 sel = 0;
        for (u64 rest = sm[k].S; rest; rest &= rest - 1) {
            const int p = __ffsll((long long)rest) - 1;
            const int times = __popcll(d.g.path_smask[sm[k].p0 + p] & pl.common);
            if (times == 0) continue;
            if (sets_count(sm[k], p) != 1) { simple = false; break; }
            tot += U[sm[k].p0 + p] * times;
            sel |= 1ull << p;
        }
        if (!simple) {
            SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
            code[k] = SFB_P2_MULTI_ALIGN;
            continue;
        }
        SFB_COUNT(sc, SFB_C_ASSIGNED_M, 1);
        code[k] = SFB_P2_ASSIGNED;
        pl.mode[k] = kStrain;
        pl.sel[k] = sel;
        pl.total[k] = tot;
        pl.n[k] = nn;
    }
}
__device__ __forceinline__ int sets_p2_need(const SetMate& s, int mode, u64 sel) {
    if (mode == kNone) return 0;
    if (mode == kSingle) return s.nt;
    return __popcll(s.m[0] & sel) + __popcll(s.m[1] & sel) + __popcll(s.m[2] & sel) + __popcll(s.m[3] & sel);
}
// The shares of one mate (json2csv.py:1155-1158, :1207-1210, :1246-1268): emit(alignment, reversed, path id, len(sequence),
// ratio, times) for every tuple that receives one.  A share names its path, not its slot: the global-memory kernels write
// work items (kAppPath) and k_pass2_scatter, one thread per item with every lane busy, looks the slot up next to the
// walk it reads anyway; the shared-memory kernel (tilesets.cu) adds it to its bins at once.  One loop over the tuples of
// the mate's four sets, so that lanes whose tuples sit in different sets (forward for one read, reverse for the next) run
// together.
constexpr uint32_t kAppPath = 0x80000000u;      // App.code = kAppPath | SFB_MATCH_REV? | path id
// every tuple of the mate whose path lies in `sel`, in the reference's order: f(set index q = 2 * alignment + reversed, path id)
template <class F>
__device__ __forceinline__ void sets_for_tuples(const SetMate& s, u64 sel, F&& f) {
    int q = 0;
    u64 cur = s.m[0] & sel;
    for (;;) {
        while (!cur && q < 3) {
            ++q;
            cur = (q == 1 ? s.m[1] : q == 2 ? s.m[2] : s.m[3]) & sel;
        }
        if (!cur) break;
        const int pid = s.p0 + __ffsll((long long)cur) - 1;
        cur &= cur - 1;
        f(q, pid);
    }
}
// single-end style: sum of the unique counts over the tuples of each alignment (json2csv.py:1240-1245)
__device__ __forceinline__ void sets_single_totals(const SetMate& s, const long long* U, long long& tot0, long long& tot1) {
    tot0 = tot1 = 0;
    for (u64 rest = s.m[0]; rest; rest &= rest - 1) tot0 += U[s.p0 + __ffsll((long long)rest) - 1];
    for (u64 rest = s.m[1]; rest; rest &= rest - 1) tot0 += U[s.p0 + __ffsll((long long)rest) - 1];
    for (u64 rest = s.m[2]; rest; rest &= rest - 1) tot1 += U[s.p0 + __ffsll((long long)rest) - 1];
    for (u64 rest = s.m[3]; rest; rest &= rest - 1) tot1 += U[s.p0 + __ffsll((long long)rest) - 1];
}
template <class Emit>
__device__ __forceinline__ void sets_p2_emit(const Dev& d, BlockCounters& sc, const Mate& mt, const SetMate& s, const SetPlan& pl,
                                             int k, const long long* U, Emit& emit) {
    const int mode = pl.mode[k];
    if (mode == kNone) return;
    long long tot0 = 0, tot1 = 0;
    if (mode == kSingle) sets_single_totals(s, U, tot0, tot1);     // every alignment separately: its own total and count
    const int cnt0 = __popcll(s.m[0]) + __popcll(s.m[1]), cnt1 = __popcll(s.m[2]) + __popcll(s.m[3]);
    sets_for_tuples(s, mode == kSingle ? ~0ull : pl.sel[k], [&](int q, int pid) {
        const i64 a = mt.a0 + (q >> 1);
        const long long u = U[pid];
        double wgt;
        int times = 1;
        if (mode == kInter) {
            wgt = ratio_of(u, pl.total[k], pl.n[k]);
        } else if (mode == kStrain) {
            times = __popcll(d.g.path_smask[pid] & pl.common);
            wgt = ratio_of(u, pl.total[k], pl.n[k]);
        } else {
            hist_add(d, sc, pid, a);                                                   // json2csv.py:1246-1252
            wgt = (q >> 1) ? ratio_of(u, tot1, cnt1) : ratio_of(u, tot0, cnt0);
        }
        emit(a, (q & 1) != 0, pid, mt.seq_len, wgt, times);
    });
}

}  // namespace sfb
