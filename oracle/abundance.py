"""Oracle (test infrastructure): the abundance pass on decoded arrays.

Restates, group by group, what ``parse_vgmpmap`` does once alignments are
decoded (reference json2csv.py:741-1279), together with
``Pangenome.get_matching_path`` (:341-363) and ``Pangenome.fill_abund_dist``
(:321-339).  Input: the Graph and Reads interchange dicts produced by
``oracle.graph`` / ``oracle.decode`` (or by the synthetic generator).

The class is deliberately sequential Python: it is the executable spec the
CUDA kernels are compared with, and the CPU baseline ``bench.py`` times.
"""
import numpy as np

# Per-mate pass-1 leaf codes (one nibble per mate in grp_code; see include/sfb200.h).
ABSENT, FILTERED, UNASSIGNED, UNASSIGNED_BOOK, DEFER, ASSIGN_SAME, ASSIGN_DIFF, \
    MULTI_ALIGN, INCONSIST, SE_COUNTED = range(10)
# Per-mate pass-2 outcome codes (grp_code2).
P2_NONE, P2_ASSIGNED, P2_MULTI_ALIGN, P2_SE_ASSIGNED, P2_UNASSIGNED, P2_INCONSIST = range(6)

COUNTERS = (
    "NB_READS", "PAIRS", "SINGLES", "FILTERED", "UNASSIGNED", "ASSIGNED_U", "ASSIGNED_M",
    "INCONSIST", "INCONSIST_M", "MULTI_ALIGN", "SE", "SAME_CP", "DIFF_CP", "ONE_NOALIGN",
    "ONE_NOCP", "AMBIG_CP_RESOLVED", "AMBIG_RESOLVED", "AMBIG_REDUCED", "SE_M", "SECOND_PASS",
)


class _Mate:
    __slots__ = ("m", "seq_len", "alns", "cm", "by_path")

    def __init__(self, m, seq_len, alns):
        self.m = m              # mate slot 0/1 inside the group
        self.seq_len = seq_len
        self.alns = alns        # global alignment indices
        self.cm = None          # colored_match: list of (pid, pos, local aln idx)
        self.by_path = None     # pid -> [local aln idx per match tuple]


class AbundanceOracle:
    @staticmethod
    def build_index(graph):
        """node -> slots where it occurs, ascending (== ascending pid, then position); the array form of
        Node.traversed_path + the position scan of json2csv.py:352-357.  Reusable across read shards."""
        path_nodes, path_off = graph["path_nodes"], graph["path_off"]
        P = len(path_off) - 1
        order = np.argsort(path_nodes, kind="stable")
        cnt = np.bincount(path_nodes, minlength=len(graph["node_len"]))
        return dict(occ_off=np.concatenate(([0], np.cumsum(cnt))).astype(np.int64), occ_slot=order.astype(np.int64),
                    slot_pid=np.repeat(np.arange(P, dtype=np.int64), np.diff(path_off)))

    def __init__(self, graph, reads, index=None):
        self.g = graph
        self.r = reads
        g = graph
        P = len(g["path_off"]) - 1
        T = int(g["path_off"][-1])
        S = len(g["strain_names"])
        self.P, self.T, self.S = P, T, S
        self.path_off = g["path_off"]
        self.path_nodes = g["path_nodes"]
        index = self.build_index(g) if index is None else index
        self.occ_off, self.occ_slot, self.slot_pid = index["occ_off"], index["occ_slot"], index["slot_pid"]
        # accumulators
        self.uniq_reads = [0] * P
        self.uniq_norm = [0] * P
        self.mult_reads = [0] * P
        self.mult_norm = [0] * P
        self.uniq_abund = np.zeros(T, np.float64)
        self.mult_abund = np.zeros(T, np.float64)
        self.hist = np.zeros((5, S, 101), np.int64)
        self.C = {k: 0 for k in COUNTERS}
        G = len(reads["grp_nmates"])
        self.grp_code = np.zeros(G, np.uint8)
        self.grp_code2 = np.zeros(G, np.uint8)
        self.aln_assign = np.full(len(reads["aln_read_len"]), -1, np.int64)  # plain slot of start
        self.book = []           # (group, mate slot, cluster id) in read order
        self.deferred = []       # group ids, in read order
        if "walk_cov" in reads:
            self.walk_cov = reads["walk_cov"]
        else:
            al = reads["walk_alen"].astype(np.int64)
            neg = al < 0
            cov = np.where(neg, 0, al & 0xFFFF) / np.where(neg, 1, al >> 16)      # aligned bases / node length
            cov[neg] = reads["exc_cov"][-al[neg] - 1]
            self.walk_cov = cov

    # ---- helpers -------------------------------------------------------
    def _walk(self, a):
        o = self.r["aln_walk_off"]
        return self.r["walk_node"][o[a]:o[a + 1]]

    def _cov(self, a):
        o = self.r["aln_walk_off"]
        return self.walk_cov[o[a]:o[a + 1]]

    def _mates(self, g):
        out = []
        off = self.r["grp_aln_off"]
        for m in range(int(self.r["grp_nmates"][g])):
            k = 2 * g + m
            out.append(_Mate(m, int(self.r["mate_seq_len"][k]), list(range(int(off[k]), int(off[k + 1])))))
        return out

    def _match_one(self, X):
        """get_matching_path (json2csv.py:341-363)."""
        res = []
        n0 = int(X[0])
        L = len(X)
        for s in self.occ_slot[self.occ_off[n0]:self.occ_off[n0 + 1]]:
            pid = int(self.slot_pid[s])
            if s + L <= self.path_off[pid + 1] and np.array_equal(self.path_nodes[s:s + L], X):
                res.append((pid, int(s - self.path_off[pid])))
        return res

    def match(self, a):
        W = self._walk(a)
        res = self._match_one(W)
        if len(W) > 1:
            res = res + self._match_one(W[::-1])
        return res

    def _color(self, mate):
        mate.cm, mate.by_path = [], {}
        for i, a in enumerate(mate.alns):
            for pid, pos in self.match(a):
                mate.cm.append((pid, pos, i))
                mate.by_path.setdefault(pid, []).append(i)

    def _strains(self, pid):
        o = self.g["pstr_off"]
        return [int(s) for s in self.g["pstr_id"][o[pid]:o[pid + 1]]]

    def _first_cluster(self, a):
        """Cluster of the first path through the walk's first node (:875,901,1036)."""
        n0 = int(self._walk(a)[0])
        if self.occ_off[n0] == self.occ_off[n0 + 1]:
            raise IndexError("list index out of range")
        pid = int(self.slot_pid[self.occ_slot[self.occ_off[n0]]])
        c = int(self.g["path_cluster"][pid])
        return None if c < 0 else c

    def _book(self, g, mate, cluster):
        if cluster is None:
            raise TypeError("list indices must be integers or slices, not NoneType")
        self.book.append((g, mate.m, cluster))

    def _hist(self, pid, a):
        st = self._strains(pid)
        if len(st) == 1:
            for k in range(5):
                b = int(self.r["aln_bins"][a, k])
                if b > 100:
                    raise KeyError(float(self.r["aln_stats"][a, k]) if "aln_stats" in self.r else b)
                self.hist[k, st[0], b] += 1

    def _fill_unique(self, pid, pos, a, seq_len):
        """fill_abund_dist (json2csv.py:321-339)."""
        self.uniq_reads[pid] += 1
        self.uniq_norm[pid] += seq_len / int(self.g["path_seq_len"][pid])
        base = int(self.path_off[pid]) + pos
        cov = self._cov(a)
        for i in range(len(cov)):
            self.uniq_abund[base + i] += cov[i]
        self.aln_assign[a] = base
        self._hist(pid, a)

    def _apply_multi(self, pid, pos, a, seq_len, ratio):
        """json2csv.py:1155-1158 (and :1207-1210, :1265-1268)."""
        self.mult_reads[pid] += ratio
        self.mult_norm[pid] += ratio * seq_len / int(self.g["path_seq_len"][pid])
        base = int(self.path_off[pid]) + pos
        cov = self._cov(a)
        for i in range(len(cov)):
            self.mult_abund[base + i] += cov[i] * ratio

    def _set_code(self, g, m, code):
        self.grp_code[g] |= np.uint8(code << (4 * m))

    def _set_code2(self, g, m, code):
        self.grp_code2[g] |= np.uint8(code << (4 * m))

    def _reduce_by_strain(self, mates):
        """Strain-intersection candidate reduction (:964-986 / :1166-1183).
        Returns None when the strain intersection is empty, else per-mate new lists."""
        by_strain, flat = [], []
        for mate in mates:
            d, f = {}, []
            for i, (pid, _, _) in enumerate(mate.cm):
                for s in self._strains(pid):
                    f.append(s)
                    d.setdefault(s, []).append(i)
            by_strain.append(d)
            flat.append(f)
        common = list(set(flat[0]) & set(flat[1]))
        if not common:
            return None
        return [[mate.cm[i] for s in common for i in d[s]] for mate, d in zip(mates, by_strain)]

    # ---- pass 1 --------------------------------------------------------
    def pass1_group(self, g):
        C = self.C
        mates = self._mates(g)
        C["NB_READS"] += len(mates)
        if len(mates) == 2:
            C["PAIRS"] += 1
            e = [len(m.alns) == 0 for m in mates]
            if e[0] and e[1]:
                C["FILTERED"] += 2
                self._set_code(g, 0, FILTERED)
                self._set_code(g, 1, FILTERED)
                return
            if e[0] or e[1]:
                C["FILTERED"] += 1
                C["ONE_NOALIGN"] += 1
                gone = 0 if e[0] else 1
                self._set_code(g, gone, FILTERED)
                mates = [mates[1 - gone]]
            else:
                for mate in mates:
                    self._color(mate)
                z = [len(m.cm) == 0 for m in mates]
                if z[0] and z[1]:
                    C["UNASSIGNED"] += 2
                    cl = [[self._first_cluster(a) for a in m.alns] for m in mates]
                    both = set(cl[0]) & set(cl[1])
                    booked = [False, False]
                    if len(both) <= 1:
                        if len(both) == 1:
                            cl = [[c for c in lst if c in both] for lst in cl]
                        for k, mate in enumerate(mates):
                            if len(cl[k]) == 1:
                                self._book(g, mate, cl[k][0])
                                booked[k] = True
                    for k in range(2):
                        self._set_code(g, k, UNASSIGNED_BOOK if booked[k] else UNASSIGNED)
                    return
                if z[0] or z[1]:
                    C["UNASSIGNED"] += 1
                    C["ONE_NOCP"] += 1
                    gone = 0 if z[0] else 1
                    cl = [self._first_cluster(a) for a in mates[gone].alns]
                    if len(cl) == 1:
                        self._book(g, mates[gone], cl[0])
                        self._set_code(g, gone, UNASSIGNED_BOOK)
                    else:
                        self._set_code(g, gone, UNASSIGNED)
                    mates = [mates[1 - gone]]
                else:
                    ids = [[t[0] for t in m.cm] for m in mates]
                    inter = list(set(ids[0]) & set(ids[1]))
                    if len(inter) > 1:
                        self.deferred.append(g)
                        C["SECOND_PASS"] += 2
                        self._set_code(g, 0, DEFER)
                        self._set_code(g, 1, DEFER)
                    elif len(inter) == 1:
                        p = inter[0]
                        for k, mate in enumerate(mates):
                            if ids[k].count(p) > 1:
                                C["MULTI_ALIGN"] += 1
                                self._set_code(g, k, MULTI_ALIGN)
                            else:
                                t = mate.cm[ids[k].index(p)]
                                C["ASSIGNED_U"] += 1
                                C["SAME_CP"] += 1
                                self._fill_unique(t[0], t[1], mate.alns[mate.by_path[p][0]], mate.seq_len)
                                if 1 < len(mate.cm):
                                    C["AMBIG_CP_RESOLVED"] += 1
                                self._set_code(g, k, ASSIGN_SAME)
                    else:
                        new = self._reduce_by_strain(mates)
                        if new is None:
                            C["INCONSIST"] += 2
                            self._set_code(g, 0, INCONSIST)
                            self._set_code(g, 1, INCONSIST)
                        else:
                            dfr = False
                            for k, mate in enumerate(mates):
                                if len(new[k]) < len(mate.cm):
                                    C["AMBIG_RESOLVED" if len(new[k]) == 1 else "AMBIG_REDUCED"] += 1
                                if len(new[k]) > 1:
                                    dfr = True
                                    C["SECOND_PASS"] += 1
                                    self._set_code(g, k, DEFER)
                                else:
                                    t = new[k][0]
                                    who = mate.by_path[t[0]]
                                    if len(who) == 1:
                                        C["ASSIGNED_U"] += 1
                                        C["DIFF_CP"] += 1
                                        self._fill_unique(t[0], t[1], mate.alns[who[0]], mate.seq_len)
                                        self._set_code(g, k, ASSIGN_DIFF)
                                    else:
                                        C["MULTI_ALIGN"] += 1
                                        self._set_code(g, k, MULTI_ALIGN)
                            if dfr:
                                self.deferred.append(g)
                    return
        # single-end style (:1003-1053)
        mate = mates[0]
        C["SINGLES"] += 1
        if not mate.alns:
            C["FILTERED"] += 1
            self._set_code(g, mate.m, FILTERED)
        elif len(mate.alns) > 1:
            self.deferred.append(g)
            C["SECOND_PASS"] += 1
            self._set_code(g, mate.m, DEFER)
        else:
            found = self.match(mate.alns[0])
            if len(found) > 1:
                self.deferred.append(g)
                C["SECOND_PASS"] += 1
                self._set_code(g, mate.m, DEFER)
            elif not found:
                C["UNASSIGNED"] += 1
                self._book(g, mate, self._first_cluster(mate.alns[0]))
                self._set_code(g, mate.m, UNASSIGNED_BOOK)
            else:
                C["ASSIGNED_U"] += 1      # counted, never accumulated (:1047-1053, Q1)
                C["SE"] += 1
                self._set_code(g, mate.m, SE_COUNTED)

    # ---- pass 2 --------------------------------------------------------
    def _ratio(self, pid, total, n):
        return 1 / n if total == 0 else self.uniq_reads_ref[pid] / float(total)

    def pass2_group(self, g):
        C = self.C
        U = self.uniq_reads_ref
        mates = self._mates(g)
        if len(mates) == 2:
            e = [len(m.alns) == 0 for m in mates]
            if e[0] or e[1]:
                mates = [mates[1]] if e[0] else [mates[0]]
            else:
                for mate in mates:
                    self._color(mate)
                z = [len(m.cm) == 0 for m in mates]
                if z[0] or z[1]:
                    mates = [mates[1]] if z[0] else [mates[0]]
                else:
                    ids = [[t[0] for t in m.cm] for m in mates]
                    inter = list(set(ids[0]) & set(ids[1]))
                    if len(inter) >= 1:
                        for mate in mates:
                            if all(len(mate.by_path[p]) == 1 for p in inter):
                                C["ASSIGNED_M"] += 1
                                total = sum(U[p] for p in inter)
                                for pid, pos, _ in mate.cm:
                                    if pid in inter:
                                        a = mate.alns[mate.by_path[pid][0]]
                                        self._apply_multi(pid, pos, a, mate.seq_len,
                                                          self._ratio(pid, total, len(inter)))
                                self._set_code2(g, mate.m, P2_ASSIGNED)
                            else:
                                C["MULTI_ALIGN"] += 1
                                self._set_code2(g, mate.m, P2_MULTI_ALIGN)
                    else:
                        new = self._reduce_by_strain(mates)
                        if new is None:
                            one = len(mates[0].cm) == 1 or len(mates[1].cm) == 1
                            C["INCONSIST_M"] += 1 if one else 2
                            self._set_code2(g, 0, P2_INCONSIST)
                            self._set_code2(g, 1, P2_INCONSIST)
                        else:
                            for k, mate in enumerate(mates):
                                if len(new[k]) > 1:
                                    if all(len(mate.by_path[t[0]]) == 1 for t in new[k]):
                                        C["ASSIGNED_M"] += 1
                                        total = sum(U[t[0]] for t in new[k])
                                        for pid, pos, _ in new[k]:
                                            a = mate.alns[mate.by_path[pid][0]]
                                            self._apply_multi(pid, pos, a, mate.seq_len,
                                                              self._ratio(pid, total, len(new[k])))
                                        self._set_code2(g, mate.m, P2_ASSIGNED)
                                    else:
                                        C["MULTI_ALIGN"] += 1
                                        self._set_code2(g, mate.m, P2_MULTI_ALIGN)
        if len(mates) == 1:
            mate = mates[0]
            hit = False
            for a in mate.alns:
                found = self.match(a)
                if found:
                    hit = True
                total = 0
                for pid, _ in found:
                    total += U[pid]
                    self._hist(pid, a)
                for pid, pos in found:
                    self._apply_multi(pid, pos, a, mate.seq_len, self._ratio(pid, total, len(found)))
            if hit:
                C["ASSIGNED_M"] += 1
                C["SE_M"] += 1
                self._set_code2(g, mate.m, P2_SE_ASSIGNED)
            else:
                C["UNASSIGNED"] += 1
                self._set_code2(g, mate.m, P2_UNASSIGNED)

    # ---- drivers -------------------------------------------------------
    def run_pass1(self, groups=None):
        G = len(self.r["grp_nmates"])
        for g in (range(G) if groups is None else groups):
            self.pass1_group(g)
        return self

    def run_pass2(self, uniq_reads=None):
        """``uniq_reads``: the global per-path unique counts to use for the
        ratios (defaults to this instance's; pass the cross-shard sum to model
        the multi-GPU exchange)."""
        self.uniq_reads_ref = self.uniq_reads if uniq_reads is None else uniq_reads
        for g in self.deferred:
            self.pass2_group(g)
        return self

    def run(self):
        return self.run_pass1().run_pass2()
