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
