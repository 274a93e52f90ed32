#!/usr/bin/env python3
"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of the path on a
6000-row corpus — K0 mask, K1 + K1F scans, K2 tcgen05 GEMM (resident) and K2T (query-tiled, in place and over the row selection's copy), K3 (block x query), K3M
(MaxScore stages and per-segment), K3D (delta rows), select, K4 fusion, ingest, index build, deletes — checked against
the oracle so that a sanitizer-clean run is also a correct one.  Run by tools/gpu_sanitize.sh."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import _coded, _data                                  # noqa: E402
from _parity import assert_same_ranking               # noqa: E402
from oracle import oracle_c                           # noqa: E402
from voitta_rag_b200 import engine                    # noqa: E402

n, dim = 6000, 64
corpus = _data.make_corpus(seed=5, n=n, dim=dim, vocab=800)
coded = _coded.code_corpus(corpus)
queries = _data.make_queries(seed=6, corpus=corpus, nq=300)
n0 = 5600
csr = coded["csr"]
sl = lambda lo, hi: (csr[0][lo:hi + 1] - csr[0][lo], csr[1][csr[0][lo]:csr[0][hi]], csr[2][csr[0][lo]:csr[0][hi]])
ix = engine.Index(dim)
ix.upsert(coded["dense"][:n0], sl(0, n0), coded["scope"][:n0], coded["created"][:n0], coded["modified"][:n0])
folders = [f for f, _ in coded["scope_list"]]
flt = (_coded.scope_bits(coded["scope_list"], include=folders[:6]), 2, 1450000000, _coded.TS_MAX)
alive = np.ones(n, np.uint8)


def check(B, rows_now, what, **opts):
    for k, v in opts.items():
        ix.set_option(k, v)
    Q = np.stack([q for q, _ in queries[:B]]); SP = [s for _, s in queries[:B]]
    fo = np.zeros(B, np.int32)
    cc = oracle_c.CorpusC(coded["dense"][:rows_now], sl(0, rows_now), coded["scope"][:rows_now], coded["created"][:rows_now],
                          coded["modified"][:rows_now], alive[:rows_now])
    got = ix.search_batch(Q, SP, [engine.Filter(*flt)], fo, limit=10, fusion="rrf", branches=True)
    want = cc.search_batch(Q, SP, [flt], fo, limit=10, fusion=2)
    for i in range(B):
        ws = [(int(want["sparse_rows"][i, j]), float(want["sparse_scores"][i, j])) for j in range(want["sparse_counts"][i])]
        assert_same_ranking(got.branch(i, "sparse"), ws, rel_tol=0.0, what=f"{what} sparse q{i}")
        wd = [(int(want["dense_rows"][i, j]), float(want["dense_scores"][i, j])) for j in range(want["dense_counts"][i])]
        assert_same_ranking(got.branch(i, "dense"), wd, rel_tol=1e-3, abs_tol=1e-3, what=f"{what} dense q{i}")
    for k in opts:
        ix.set_option(k, {"dense_path": 0, "k1f": 1, "sparse_ms": 1, "ms_staged": 1, "safe_mode": 0}[k])
    print("ok", what, flush=True)


check(1, n0, "B=1 K1F + K3M stages")
check(1, n0, "B=1 K1 segments + K3M per segment", k1f=0, ms_staged=0)
check(3, n0, "B=3 K2 resident", dense_path=2)
check(3, n0, "B=3 K3 only", sparse_ms=0)
check(3, n0, "B=3 safe mode", safe_mode=1)
check(300, n0, "B=300 K2T")
ix.set_option("dense_compact_min_rows", 1024)         # the row selection normally needs 262144-row segments ...
ix.set_option("dense_compact", 100)                   # ... and a filter selective enough to pay for the copy (here: always)
check(300, n0, "B=300 K2T over the compacted copy of the passing rows (row selection)")
st = ix.stats()
assert st["last_sel_rows"] > 0 and st["last_sel_used"] == 1, st
ix.set_option("dense_compact_min_rows", 1 << 18)
ix.set_option("dense_compact", -1)
ix.set_option("ms_max_terms", 4)                      # queries of more than 4 terms become "long"
check(6, n0, "B=6 long queries on K3")
ix.set_option("sparse_mh", 1)                         # ... and on K3H (off by default)
check(6, n0, "B=6 long queries on K3H")
ix.set_option("sparse_mh", 0)
ix.set_option("ms_max_terms", 16)
ix.upsert(coded["dense"][n0:n], sl(n0, n), coded["scope"][n0:n], coded["created"][n0:n], coded["modified"][n0:n])
dead = np.arange(100, 400, 7)
ix.delete_rows(dead.astype(np.uint64)); alive[dead] = 0
check(2, n, "delta rows + deletes (K3D)")
ix.optimize()
check(2, n, "after merge")
ix.close()
print("sanitize workload done")
