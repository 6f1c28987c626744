"""Oracle (test infrastructure): the CPU port spread over host cores, for bench.py's CPU baseline.

The reference is a single Python thread.  To give the CPU side every core, read groups are sharded over
worker processes exactly as the multi-GPU path shards them: pass 1 per shard, sum
``uniq_reads`` over shards, pass 2 per shard, sum the accumulators.  Workers are forked once and keep the
graph index; only the abundance pass is timed by the caller.
"""
import multiprocessing as mp
import os

import numpy as np

from .abundance import AbundanceOracle

_G = {}


def _slice_reads(r, g0, g1):
    a0, a1 = int(r["grp_aln_off"][2 * g0]), int(r["grp_aln_off"][2 * g1])
    w0, w1 = int(r["aln_walk_off"][a0]), int(r["aln_walk_off"][a1])
    out = dict(
        grp_nmates=r["grp_nmates"][g0:g1], grp_aln_off=r["grp_aln_off"][2 * g0:2 * g1 + 1] - a0,
        mate_seq_len=r["mate_seq_len"][2 * g0:2 * g1], aln_walk_off=r["aln_walk_off"][a0:a1 + 1] - w0,
        aln_read_len=r["aln_read_len"][a0:a1], aln_bins=r["aln_bins"][a0:a1], walk_node=r["walk_node"][w0:w1],
        walk_alen=r["walk_alen"][w0:w1], exc_cov=r["exc_cov"])
    if "walk_cov" in r:
        out["walk_cov"] = r["walk_cov"][w0:w1]
    return out


def _pass1(args):
    g0, g1 = args
    orc = AbundanceOracle(_G["graph"], _slice_reads(_G["reads"], g0, g1), index=_G["index"])
    orc.run_pass1()
    _G["orc"] = orc
    return np.asarray(orc.uniq_reads, np.int64)


def _sparse(a):
    idx = np.flatnonzero(a)
    return idx, a[idx]


def _pass2(uniq):
    """The slot arrays go back as (index, value) lists: a shard touches a small part of the graph, and pickling the
    dense arrays of every worker through a pipe used to cost more than the pass itself."""
    orc = _G.pop("orc")
    orc.run_pass2(uniq_reads=[int(v) for v in uniq])
    return dict(uniq_reads=np.asarray(orc.uniq_reads, np.int64), uniq_norm=np.asarray(orc.uniq_norm, np.float64),
                mult_reads=np.asarray(orc.mult_reads, np.float64), mult_norm=np.asarray(orc.mult_norm, np.float64),
                mult_hits=np.array([not isinstance(v, int) for v in orc.mult_reads]),
                uniq_abund=_sparse(orc.uniq_abund), mult_abund=_sparse(orc.mult_abund), hist=orc.hist, C=orc.C,
                n_slots=len(orc.uniq_abund), n_deferred=len(orc.deferred))


class ParallelOracle:
    """One worker process per shard; every worker handles exactly one shard per pass (pinned by using
    ``cores`` single-worker pools), so that pass 2 finds the state pass 1 left behind."""

    def __init__(self, graph, reads, cores=None):
        self.cores = max(1, cores or len(os.sched_getaffinity(0)))
        _G.update(graph=graph, reads=reads, index=AbundanceOracle.build_index(graph))
        ctx = mp.get_context("fork")
        self.pools = [ctx.Pool(1) for _ in range(self.cores)]
        G = len(reads["grp_nmates"])
        cuts = [G * k // self.cores for k in range(self.cores + 1)]
        self.shards = list(zip(cuts[:-1], cuts[1:]))

    def run(self):
        parts = [p.apply_async(_pass1, (s,)) for p, s in zip(self.pools, self.shards)]
        uniq = np.sum([x.get() for x in parts], axis=0)
        parts = [p.apply_async(_pass2, (uniq,)) for p in self.pools]
        res = [x.get() for x in parts]
        out = {k: np.sum([r[k] for r in res], axis=0) for k in ("uniq_reads", "uniq_norm", "mult_reads", "mult_norm", "hist")}
        out["mult_hits"] = np.any([r["mult_hits"] for r in res], axis=0)
        for k in ("uniq_abund", "mult_abund"):
            dense = np.zeros(res[0]["n_slots"], np.float64)
            for r in res:
                idx, val = r[k]
                dense[idx] += val            # the indices of one shard are distinct
            out[k] = dense
        out["C"] = {k: sum(r["C"][k] for r in res) for k in res[0]["C"]}
        out["n_deferred"] = sum(r["n_deferred"] for r in res)
        return out

    def close(self):
        for p in self.pools:
            p.terminate()
        _G.clear()
