"""The drop-in VectorStoreService on the real B200 backend replays the scenario recorded from the
REFERENCE's own vector_store.py (tests/golden/voitta_cases.json)."""
import json
from pathlib import Path

import pytest

from golden import make_golden as G
from _parity import assert_same_ranking

pytestmark = pytest.mark.gpu

GOLD = json.loads((Path(__file__).parent / "golden" / "voitta_cases.json").read_text())


@pytest.mark.parametrize("dense_path", [0, 2])
def test_reference_scenario_on_b200(monkeypatch, dense_path):
    from voitta_rag_b200 import vector_store as VS
    name = f"gpu_golden_{dense_path}"
    monkeypatch.setenv("QDRANT_COLLECTION", name)
    monkeypatch.setenv("EMBEDDING_DIMENSION", str(GOLD["dim"]))
    VS._drop_collection(name)
    store = VS.VectorStoreService()
    store.client.set_option("dense_path", dense_path)
    corpus, queries = G.build_inputs()
    outs = G.replay(store, GOLD["ops"], corpus, queries, VS.ChunkMetadata, {})
    # K1 (fp32 query) agrees to summation error; K2 feeds the query as bf16 hi+lo halves (B = 1 fits
    # one pass), so it agrees to ~2e-5 as well
    tol = 1e-5 if dense_path == 0 else 2e-5
    assert len(outs) == len(GOLD["outs"])
    for i, (op, got, want) in enumerate(zip(GOLD["ops"], outs, GOLD["outs"])):
        what = f"op {i} {op}"
        if op["op"] == "search":
            if dense_path == 2 and op["kw"].get("sparse") not in (False, "empty"):
                # fused scores are min-max normalised: an error e of the k'-th dense score moves every
                # normalised score by ~e/spread
                assert len(got) == len(want), what
                assert_same_ranking(got, want, rel_tol=5e-4, abs_tol=5e-4, what=what)
            else:
                assert_same_ranking(got, want, rel_tol=tol, abs_tol=tol, what=what)
                for g, w in zip(got, want):
                    if g[0] == w[0]:
                        assert g[2:] == w[2:], what
        else:
            assert json.loads(json.dumps(got)) == want, what
    assert store.client.stats()["searches"] >= 40
    VS._drop_collection(name)


def test_hand_derived_known_answers_on_b200():
    """tests/golden/known_answers.json (worked by hand from the published formulas) through the C ABI:
    IDF + sparse dot + exclusion of rows without a shared index, cosine with a zero row, range on a missing key."""
    import numpy as np
    from voitta_rag_b200 import engine
    KA = json.loads((Path(__file__).parent / "golden" / "known_answers.json").read_text())
    close = lambda a, b: abs(a - b) <= 2e-7 * max(1.0, abs(a), abs(b))
    # sparse
    s = KA["sparse"]
    tid = s["term_ids"]
    n = len(s["docs"])
    indptr = np.zeros(n + 1, np.int64)
    terms, vals = [], []
    for i, d in enumerate(s["docs"]):
        pairs = sorted((tid[t], float(v)) for t, v in d.items())
        terms += [p[0] for p in pairs]; vals += [p[1] for p in pairs]
        indptr[i + 1] = len(terms)
    ix = engine.Index(2)
    ix.upsert(np.tile(np.array([[1.0, 0.0]], np.float32), (n, 1)), (indptr, np.array(terms, np.uint32), np.array(vals, np.float32)))
    df, n_live = ix.term_stats(np.array([tid[t] for t in "ABCDE"], np.uint32))
    assert n_live == 5 and [int(x) for x in df] == [1, 2, 5, 3, 0]
    q = ([tid[t] for t in s["query"]], [float(v) for v in s["query"].values()])
    got = ix.search_batch(np.array([[1.0, 0.0]], np.float32), [q], limit=5, fusion="rrf", branches=True)
    br = got.branch(0, "sparse")
    assert [r for r, _ in br] == [r for r, _ in s["ranking"]]
    for (_, sc), (_, want) in zip(br, s["ranking"]):
        assert close(sc, want), (sc, want)
    ix.close()
    # cosine with a zero row, both dense kernels
    c = KA["cosine"]
    ix = engine.Index(2)
    ix.upsert(np.array(c["rows"], np.float32))
    for path in (1, 2):
        ix.set_option("dense_path", path)
        got = ix.search_batch(np.array([c["query"], c["query"]], np.float32), None, limit=4, fusion="dense", branches=True)
        sc = dict(got.branch(0, "dense"))
        for r, want in enumerate(c["scores"]):
            assert abs(sc[r] - want) <= 1e-6, (path, r, sc[r], want)
        assert [r for r, _ in got.branch(0, "dense")][0] == 0 and [r for r, _ in got.branch(0, "dense")][-1] == 2
    ix.close()
    # range on a missing key
    rk = KA["range_missing_key"]
    miss = engine.TS_MISSING
    mod = np.array([miss if v is None else v for v in rk["modified"]], np.int64)
    cre = np.array([miss if v is None else v for v in rk["created"]], np.int64)
    ix = engine.Index(2)
    ix.upsert(np.tile(np.array([[1.0, 0.0]], np.float32), (4, 1)), None, None, cre, mod)
    for case in rk["cases"]:
        field = engine.TS_CREATED if case["date_field"] == "created" else engine.TS_MODIFIED
        lo = engine.TS_MIN if case["date_start"] is None else case["date_start"]
        hi = engine.TS_MAX if case["date_end"] is None else case["date_end"]
        got = ix.search_batch(np.array([[1.0, 0.0]], np.float32), None, [engine.Filter(None, field, lo, hi)], np.zeros(1, np.int32),
                              limit=4, fusion="dense")
        assert sorted(int(r) for r in got.rows[0, :got.counts[0]]) == case["pass"], case["why"]
    ix.close()
