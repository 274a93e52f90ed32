"""oracle/cpu_fast.py (BLAS + scipy CSR + vectorised masks / fusion) against the two loop-shaped oracles on the same
seeded inputs: the Python restatement of qdrant-client local mode (oracle.py, small n) and its C port (oracle_c.c,
n up to 50k, as BASELINE.md §3 asks).  Sparse lists must be bit-equal (ordered float64 sums in all three), dense
lists equal within fp32 summation error, fused lists equal given the same branch lists."""
import numpy as np
import pytest

import _coded
import _data
from _parity import assert_same_ranking
from oracle import cpu_fast, oracle_c
from oracle import oracle as O

FZ = {"dense": 0, "weighted": 1, "rrf": 2}


def lists(out, i, which):
    pre = "" if which == "fused" else which + "_"
    c = out[("counts" if which == "fused" else pre + "counts")][i]
    return [(int(out[pre + "rows"][i, j]), float(out[pre + "scores"][i, j])) for j in range(c)]


def filters_for(coded):
    sl = coded["scope_list"]
    folders = [f for f, _ in sl]
    return [None,
            (_coded.scope_bits(sl, include=folders[:5]), 0, _coded.TS_MIN, _coded.TS_MAX),
            (_coded.scope_bits(sl, exclude=folders[:2], disabled=["root3"]), 0, _coded.TS_MIN, _coded.TS_MAX),
            (None, 2, 1500000000, 1650000000),
            (_coded.scope_bits(sl, include=folders[3:20]), 1, 1450000000, _coded.TS_MAX),
            (_coded.scope_bits(sl, include=[folders[-1]]), 2, 1766000000, 1767225600)]


@pytest.mark.parametrize("n,seed", [(3000, 11), (50_000, 12)])
@pytest.mark.parametrize("fusion", ["weighted", "rrf"])
def test_cpu_fast_equals_c_oracle(n, seed, fusion):
    dim = 48
    corpus = _data.make_corpus(seed=seed, n=n, dim=dim, n_index=5, per_index=7, vocab=2500)
    corpus["dense"][7] = 0.0                                # zero row: cosine 0 (norm 0 -> EPSILON)
    corpus["dense"][9] = corpus["dense"][8]                 # exact tie: lower row first
    corpus["sparse"][9] = corpus["sparse"][8]
    queries = _data.make_queries(seed=seed + 100, corpus=corpus, nq=10)
    coded = _coded.code_corpus(corpus)
    alive = np.ones(n, np.uint8)
    alive[np.random.RandomState(seed).choice(n, size=n // 50, replace=False)] = 0
    args = (coded["dense"], coded["csr"], coded["scope"], coded["created"], coded["modified"], alive)
    cc, cf = oracle_c.CorpusC(*args), cpu_fast.CorpusFast(*args)
    Q = np.stack([q for q, _ in queries])
    SP = [s for _, s in queries]
    SP[3] = None                                            # a dense-only query inside a hybrid batch
    flts = filters_for(coded)
    for f in flts:
        fl, fo = (None, None) if f is None else ([f], np.zeros(len(Q), np.int32))
        want = cc.search_batch(Q, SP, fl, fo, limit=10, fusion=FZ[fusion], sparse_weight=0.2)
        got = cf.search_batch(Q, SP, fl, fo, limit=10, fusion=FZ[fusion], sparse_weight=0.2)
        for i in range(len(Q)):
            assert lists(got, i, "sparse") == lists(want, i, "sparse"), f"sparse q{i} filter {f and f[1:]}"
            assert_same_ranking(lists(got, i, "dense"), lists(want, i, "dense"), rel_tol=1e-5, abs_tol=1e-5, what=f"dense q{i}")
            assert_same_ranking(lists(got, i, "fused"), lists(want, i, "fused"), rel_tol=1e-4, abs_tol=1e-4, what=f"fused q{i}")
    # several filters in one batch, one per query
    fl = [f for f in flts if f is not None]
    fo = (np.arange(len(Q)) % (len(fl) + 1) - 1).astype(np.int32)          # -1 = unfiltered
    want = cc.search_batch(Q, SP, fl, fo, limit=7, fusion=FZ[fusion])
    got = cf.search_batch(Q, SP, fl, fo, limit=7, fusion=FZ[fusion])
    for i in range(len(Q)):
        assert lists(got, i, "sparse") == lists(want, i, "sparse")
        assert_same_ranking(lists(got, i, "dense"), lists(want, i, "dense"), rel_tol=1e-5, abs_tol=1e-5, what=f"dense q{i}")


def test_cpu_fast_equals_python_oracle_branches_and_fusion():
    """Against oracle.py itself (the line-cited restatement): branch lists through LocalCollection, then both fusions
    fed the SAME lists must give identical ids and float64 scores."""
    n, dim = 1500, 32
    corpus = _data.make_corpus(seed=5, n=n, dim=dim, vocab=900)
    queries = _data.make_queries(seed=6, corpus=corpus, nq=6)
    coded = _coded.code_corpus(corpus)
    cf = cpu_fast.CorpusFast(coded["dense"], coded["csr"], coded["scope"], coded["created"], coded["modified"])
    coll = O.LocalCollection(dim)
    for r in range(n):
        coll.upsert(str(r), corpus["dense"][r], corpus["sparse"][r], {})
    for qi, (q, sp) in enumerate(queries):
        out = cf.search_batch(q, [sp], limit=8, fusion=1, sparse_weight=0.3)
        d = coll.query_dense(q, 24)
        s = coll.query_sparse(sp[0], sp[1], 24)
        assert [(p.row, p.score) for p in s] == lists(out, 0, "sparse"), f"q{qi} sparse"
        assert_same_ranking(lists(out, 0, "dense"), [(p.row, p.score) for p in d], rel_tol=1e-5, abs_tol=1e-5, what=f"q{qi} dense")
        # fusion on identical inputs: feed cpu_fast's own lists to the reference-shaped fusion functions
        dl = [O.ScoredPoint(str(r), sc, {}, r) for r, sc in lists(out, 0, "dense")]
        sl = [O.ScoredPoint(str(r), sc, {}, r) for r, sc in lists(out, 0, "sparse")]
        want_w = [(int(pid), sc) for pid, sc, _ in O.weighted_fusion(dl, sl, 8, 0.3)]
        assert want_w == lists(out, 0, "fused"), f"q{qi} weighted fusion"
        out_r = cf.search_batch(q, [sp], limit=8, fusion=2)
        dl = [O.ScoredPoint(str(r), sc, {}, r) for r, sc in lists(out_r, 0, "dense")]
        sl = [O.ScoredPoint(str(r), sc, {}, r) for r, sc in lists(out_r, 0, "sparse")]
        want_r = [(int(pid), sc) for pid, sc, _ in O.reciprocal_rank_fusion([dl, sl], 8)]
        assert want_r == lists(out_r, 0, "fused"), f"q{qi} rrf"
