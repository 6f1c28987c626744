// Warp-cooperative pass-2 logic for read groups with many candidate paths (more than the register path's 8
// match tuples per mate, up to 32 per mate): one warp per group, one lane per tuple.
//
//   "is this path also matched by the other mate"   loop over the other mate's tuples, one shuffle each
//   "how often does this path occur in my mate"      __match_any_sync on the path id
//   sizes of sets, sums of unique-read counts         ballots / shuffle reductions (segmented by alignment for the
//                                                     single-end style)
//   strain algebra                                    64-bit OR-reduction per mask word, popcount per lane
//
// Included by group.cu inside namespace sfb, after the shared helpers (assign_unique, hist_add, emit_app, ...).
// Larger groups (> 32 tuples in a mate) and the rare no-colored-path cases are handled by lane 0 with the generic
// serial code.
#pragma once

struct WarpMate {
    int n;          // tuples of the mate (warp-uniform)
    int pid;        // this lane's tuple (a distinct negative id when the lane has none)
    uint32_t code;
    int aloc;
    bool valid;
};

// false when the mate has more than 32 tuples (warp-uniform result)
__device__ __forceinline__ bool warp_load(const sfb_work& w, const Mate& mt, int sentinel, WarpMate& m) {
    const int lane = lane_id();
    m.pid = sentinel - lane;
    m.code = 0;
    m.aloc = 0;
    m.valid = false;
    int base = 0;
    for (int j = 0; j < mt.na; ++j) {
        const uint2 mw = match_word(w, mt.a0 + j);
        const int c = match_count(mw);
        if (base + c > 32) return false;
        if (lane >= base && lane < base + c) {
            const uint2 t = match_get(w, mw, lane - base);
            m.code = t.x;
            m.pid = (int)t.y;
            m.aloc = j;
            m.valid = true;
        }
        base += c;
    }
    m.n = base;
    return true;
}

__device__ __forceinline__ bool warp_in_other(const WarpMate& mine, const WarpMate& other) {
    bool in = false;
    for (int j = 0; j < other.n; ++j) in |= __shfl_sync(0xffffffffu, other.pid, j) == mine.pid;
    return in;
}
__device__ __forceinline__ u64 warp_or(u64 v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// times[k]: in how many common strains this lane's path occurs; nn[k]: size of mate k's reduced list (Q4)
__device__ __forceinline__ bool warp_strain_cut(const sfb_graph& g, const WarpMate& A, const WarpMate& B, int* times,
                                                int* nn) {
    times[0] = times[1] = 0;
    for (int word = 0; word < (int)g.mask_words; ++word) {
        const u64 s0 = A.valid ? g.path_smask[(size_t)A.pid * g.mask_words + word] : 0ull;
        const u64 s1 = B.valid ? g.path_smask[(size_t)B.pid * g.mask_words + word] : 0ull;
        const u64 common = warp_or(s0) & warp_or(s1);
        times[0] += __popcll(s0 & common);
        times[1] += __popcll(s1 & common);
    }
    nn[0] = warp_sum(times[0]);
    nn[1] = warp_sum(times[1]);
    return nn[0] > 0;
}

constexpr int kWarpsGroup = 4;      // warps per block of the warp-cooperative kernel

// ---- pass 2 (json2csv.py:1085-1274) ------------------------------------------------------------------------
constexpr int kStageApps = 128;     // staged work items per warp (a group adds at most 64)

__global__ void __launch_bounds__(128) k_pass2_resolve_slow(Dev d, const long long* __restrict__ U) {
    __shared__ BlockCounters sc;
    __shared__ App s_app[kWarpsGroup][kStageApps];
    block_counters_zero(sc);
    __syncthreads();
    const int wi = threadIdx.x >> 5, lane = lane_id();
    const unsigned below = (1u << lane) - 1u;
    const i64 n_slow = (i64)d.w.cursor[4];
    const i64 n_warps = (i64)gridDim.x * kWarpsGroup;
    App* apps = reinterpret_cast<App*>(d.w.app_buf);
    int n_stage = 0;                     // warp-uniform fill of s_app[wi]
    auto flush = [&]() {
        u64 base = 0;
        if (lane == 0) base = atomicAdd(d.w.cursor + 2, (u64)n_stage);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        if (base + n_stage <= (u64)d.w.app_cap) {
            for (int i = lane; i < n_stage; i += 32) apps[base + i] = s_app[wi][i];
        } else if (lane == 0) {
            SFB_COUNT(sc, SFB_C_MATCH_OVERFLOW, n_stage);
        }
        __syncwarp();
        n_stage = 0;
    };
    // one lane's share: per-path counters now, the node loop later (k_pass2_scatter)
    auto stage = [&](bool take, const Mate& mt, const WarpMate& M, double ratio, int times) {
        const unsigned sel = __ballot_sync(0xffffffffu, take);
        if (take) {
            const int pid = M.pid;
            acc_add(&d.acc.mult_reads[pid], d.acc.det_stride, ratio * (double)times);
            acc_add(&d.acc.mult_norm[pid], d.acc.det_stride,
                    (ratio * (double)mt.seq_len / (double)d.g.path_seq_len[pid]) * (double)times);
            d.acc.mult_hits[pid] = 1u;       // "touched" flag, see emit_app
            App app;
            app.aln = (int32_t)(mt.a0 + M.aloc);
            app.code = M.code;
            app.weight = ratio * (double)times;
            s_app[wi][n_stage + __popc(sel & below)] = app;
        }
        n_stage += __popc(sel);
        if (lane == 0 && sel) SFB_COUNT(sc, SFB_C_ST_P2_APPLIED, __popc(sel));
    };
    for (i64 q = (i64)blockIdx.x * kWarpsGroup + wi; q < n_slow; q += n_warps) {
        const i64 g = d.w.slow[q];
        const int nm = d.r.grp_nmates[g];
        Mate mt[2];
        mate_load(d.r, g, 0, mt[0]);
        mate_load(d.r, g, 1, mt[1]);
        int code[2] = {SFB_P2_NONE, SFB_P2_NONE};
        WarpMate A, B;
        const bool fits = warp_load(d.w, mt[0], -2, A) && warp_load(d.w, mt[1], -40, B);
        if (!fits) {
            // more than 32 tuples in a mate: lane 0, generic serial code, own claim
            if (lane == 0) {
                GlobalView gv;
                gv.init(d.w, mt);
                GlobalOps ops{U, nullptr};
                Plan pl;
                pl.mode[0] = pl.mode[1] = kNone;
                pl.napp[0] = pl.napp[1] = 0;
                int single = nm == 1 ? 0 : -1;
                if (nm == 2) {
                    if (mt[0].na == 0 || mt[1].na == 0) single = mt[0].na == 0 ? 1 : 0;
                    else if (gv.nt(0) == 0 || gv.nt(1) == 0) single = gv.nt(0) == 0 ? 1 : 0;
                    else p2_decide_pair(d, ops, sc, gv, pl, code);
                }
                if (single >= 0) {
                    pl.mode[single] = kSingle;
                    pl.napp[single] = gv.nt(single);
                    code[single] = p2_single_outcome(sc, gv.nt(single));
                }
                const unsigned need = (unsigned)(pl.napp[0] + pl.napp[1]);
                if (need) {
                    const u64 pos = atomicAdd(d.w.cursor + 2, (u64)need);
                    SFB_COUNT(sc, SFB_C_ST_P2_APPLIED, (int)need);
                    if (pos + need <= (u64)d.w.app_cap) { ops.out = apps + pos; p2_emit(d, ops, sc, gv, mt, pl); }
                    else SFB_COUNT(sc, SFB_C_MATCH_OVERFLOW, (int)need);
                }
                d.w.grp_code2[g] = (uint8_t)(code[0] | (code[1] << 4));
            }
            continue;
        }
        int single = nm == 1 ? 0 : -1;
        if (nm == 2) {
            if (mt[0].na == 0 || mt[1].na == 0) single = mt[0].na == 0 ? 1 : 0;          // json2csv.py:1088-1093
            else if (A.n == 0 || B.n == 0) single = A.n == 0 ? 1 : 0;                    // :1117-1122
        }
        if (single >= 0) {
            // every alignment separately (json2csv.py:1230-1268)
            const WarpMate& M = single ? B : A;
            const Mate& mm = mt[single];
            const long long u = M.valid ? U[M.pid] : 0;
            double ratio = 0.0;
            for (int j = 0; j < mm.na; ++j) {
                const bool mine = M.valid && M.aloc == j;
                const long long tot = warp_sum_ll(mine ? u : 0);
                const int cnt = __popc(__ballot_sync(0xffffffffu, mine));
                if (mine) ratio = ratio_of(u, tot, cnt);
            }
            if (M.valid) hist_add(d, sc, M.pid, mm.a0 + M.aloc);
            stage(M.valid, mm, M, ratio, 1);
            if (lane == 0) code[single] = p2_single_outcome(sc, M.n);
        } else {
            const bool in0 = warp_in_other(A, B), in1 = warp_in_other(B, A);
            const unsigned same0 = __match_any_sync(0xffffffffu, A.pid), same1 = __match_any_sync(0xffffffffu, B.pid);
            const bool first0 = A.valid && (__ffs(same0) - 1 == lane);
            const int n_inter = __popc(__ballot_sync(0xffffffffu, in0 && first0));
            if (n_inter >= 1) {
                const long long total = warp_sum_ll(in0 && first0 ? U[A.pid] : 0);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const WarpMate& M = k ? B : A;
                    const bool in = k ? in1 : in0;
                    const bool dup = __popc(k ? same1 : same0) > 1;
                    if (__any_sync(0xffffffffu, in && dup)) {
                        if (lane == 0) SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
                        code[k] = SFB_P2_MULTI_ALIGN;
                        continue;
                    }
                    if (lane == 0) SFB_COUNT(sc, SFB_C_ASSIGNED_M, 1);
                    code[k] = SFB_P2_ASSIGNED;
                    stage(in, mt[k], M, in ? ratio_of(U[M.pid], total, n_inter) : 0.0, 1);
                }
            } else {
                int times[2], nn[2];
                if (!warp_strain_cut(d.g, A, B, times, nn)) {
                    if (lane == 0) SFB_COUNT(sc, SFB_C_INCONSIST_M, (A.n == 1 || B.n == 1) ? 1 : 2);
                    code[0] = code[1] = SFB_P2_INCONSIST;
                } else {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const WarpMate& M = k ? B : A;
                        if (nn[k] <= 1) continue;                      // handled in pass 1 (json2csv.py:1185)
                        const bool sel = times[k] > 0;
                        const bool dup = __popc(k ? same1 : same0) > 1;
                        if (__any_sync(0xffffffffu, sel && dup)) {
                            if (lane == 0) SFB_COUNT(sc, SFB_C_MULTI_ALIGN, 1);
                            code[k] = SFB_P2_MULTI_ALIGN;
                            continue;
                        }
                        const long long u = sel ? U[M.pid] : 0;
                        const long long tot = warp_sum_ll(u * times[k]);
                        if (lane == 0) SFB_COUNT(sc, SFB_C_ASSIGNED_M, 1);
                        code[k] = SFB_P2_ASSIGNED;
                        stage(sel, mt[k], M, sel ? ratio_of(u, tot, nn[k]) : 0.0, times[k]);
                    }
                }
            }
        }
        if (lane == 0) d.w.grp_code2[g] = (uint8_t)(code[0] | (code[1] << 4));
        if (n_stage > kStageApps - 64) flush();
    }
    if (n_stage) flush();
    __syncthreads();
    block_counters_flush(sc, d.acc.counters);
}
