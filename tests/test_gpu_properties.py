"""GPU, sizes the Python oracle cannot reach: size-independent properties of the abundance pass
(conservation laws between counters, leaf codes and accumulators; idempotence)."""
import numpy as np
import pytest
import torch

from strainflair_b200 import _lib, engine, synth

pytestmark = pytest.mark.gpu


def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def cov_of(reads):
    v = reads.walk_alen.astype(np.int64)
    neg = v < 0
    cov = np.where(neg, 0, v & 0xFFFF) / np.where(neg, 1, v >> 16)
    if neg.any():
        cov[neg] = reads.exc_cov[-v[neg] - 1]
    return cov


@pytest.mark.parametrize("name,scale", [("c3", 0.04), ("c4", 0.01)])
def test_conservation_laws(name, scale):
    need_gpu()
    graph, reads = synth.make_config(name, scale=scale)
    eng = engine.AbundanceEngine(graph, tile_kernels=True)
    res = eng.run(reads)
    C = res.C
    lo, hi = res.grp_code & 15, res.grp_code >> 4
    n = lambda code: int((lo == code).sum() + (hi == code).sum())          # noqa: E731
    lo2, hi2 = res.grp_code2 & 15, res.grp_code2 >> 4
    n2 = lambda code: int((lo2 == code).sum() + (hi2 == code).sum())       # noqa: E731
    # every mate has exactly one pass-1 leaf, and the counters are the leaf counts
    assert C["NB_READS"] == int(reads.grp_nmates.sum()) == sum(n(c) for c in range(1, 10))
    assert C["PAIRS"] == int((reads.grp_nmates == 2).sum())
    assert C["FILTERED"] == n(_lib.FILTERED)
    assert C["SECOND_PASS"] == n(_lib.DEFER)
    assert C["SAME_CP"] == n(_lib.ASSIGN_SAME) and C["DIFF_CP"] == n(_lib.ASSIGN_DIFF)
    assert C["SE"] == n(_lib.SE_COUNTED) and C["INCONSIST"] == n(_lib.INCONSIST)
    assert C["ASSIGNED_U"] == C["SAME_CP"] + C["DIFF_CP"] + C["SE"]
    assert C["MULTI_ALIGN"] == n(_lib.MULTI_ALIGN) + n2(2)
    assert C["UNASSIGNED"] == n(_lib.UNASSIGNED) + n(_lib.UNASSIGNED_BOOK) + n2(4)
    assert C["ASSIGNED_M"] == n2(1) + n2(3) and C["SE_M"] == n2(3)
    assert res.n_deferred == int(((lo == _lib.DEFER) | (hi == _lib.DEFER)).sum())
    # unique reads: one count per filled assignment, and the slot sums are the assigned walks' coverages
    assert int(res.uniq_reads.sum()) == C["SAME_CP"] + C["DIFF_CP"] == int((res.aln_assign >= 0).sum())
    cov = cov_of(reads)
    per_aln = np.add.reduceat(cov, reads.aln_walk_off[:-1].clip(max=len(cov) - 1)) * (np.diff(reads.aln_walk_off) > 0)
    want = float(per_aln[res.aln_assign >= 0].sum())
    assert abs(float(res.uniq_abund.sum()) - want) <= 1e-9 * want
    L = np.diff(reads.aln_walk_off)
    assert C["ST_P1_RECORDS"] == int(L[res.aln_assign >= 0].sum())
    # the two kernel families (shared-memory tiles above, global memory here) agree: integers exactly, and -- the sums
    # being exact integer sums of the same terms in the default order-independent mode -- the fp64 arrays bit for bit
    assert eng.tiling is not None and eng.tiling.n_tiles > 1000
    dflt = engine.AbundanceEngine(graph).run(reads)                       # global-memory kernels on tile-ordered reads
    glob = engine.AbundanceEngine(graph, tiles=False).run(reads)          # ... and on reads in file order
    for name in ("grp_code", "grp_code2", "aln_assign", "uniq_reads", "hist", "mult_hits", "uniq_norm", "mult_reads",
                 "mult_norm", "uniq_abund", "mult_abund", "table"):
        assert np.array_equal(getattr(glob, name), getattr(dflt, name), equal_nan=True), name
    assert glob.C == {**res.C, **{k: glob.C[k] for k in ("ST_CAND", "ST_COMPARES")}}
    for name in ("grp_code", "grp_code2", "aln_assign", "uniq_reads", "hist", "mult_hits"):
        assert np.array_equal(getattr(glob, name), getattr(res, name)), name
    for name in ("uniq_norm", "mult_reads", "mult_norm", "uniq_abund", "mult_abund", "table"):
        assert np.array_equal(getattr(glob, name), getattr(res, name), equal_nan=True), name
    # pass 2 hands out exactly one read per (mate, alignment) unit: the shares of a unit sum to 1
    x, y = glob.aln_match[:, 0], glob.aln_match[:, 1]
    has_tuples = x != _lib.MATCH_NONE
    units = n2(1)
    for m, codes in ((0, lo2), (1, hi2)):
        for g in np.flatnonzero(codes == 3):                                  # single-end style: per alignment
            a0, a1 = int(reads.grp_aln_off[2 * g + m]), int(reads.grp_aln_off[2 * g + m + 1])
            units += int(has_tuples[a0:a1].sum())
    assert abs(float(res.mult_reads.sum()) - units) <= 1e-9 * max(1, units)
    assert int((res.mult_hits > 0).sum()) == int((res.mult_reads > 0).sum()) or (res.mult_reads[res.mult_hits > 0] >= 0).all()
    # the table is consistent with the accumulators
    plen = np.diff(graph.path_off)
    mean_u = np.add.reduceat(res.uniq_abund, graph.path_off[:-1]) / plen
    assert np.allclose(res.table[:, 3], mean_u, rtol=1e-9, atol=0)
    assert np.allclose(res.table[:, 1], res.uniq_reads + res.mult_reads, rtol=1e-12, atol=0)
    # idempotence: a second run gives the same integers and (atomic order aside) the same floats
    res2 = eng.run(reads)
    assert res2.C == res.C and np.array_equal(res2.uniq_reads, res.uniq_reads) and np.array_equal(res2.hist, res.hist)
    assert np.array_equal(res2.grp_code, res.grp_code) and np.array_equal(res2.grp_code2, res.grp_code2)
    assert np.allclose(res2.mult_abund, res.mult_abund, rtol=1e-12, atol=0)
    assert np.allclose(res2.uniq_abund, res.uniq_abund, rtol=1e-12, atol=0)
