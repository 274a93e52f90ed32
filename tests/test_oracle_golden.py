"""The oracle restatement of voitta's own logic vs outputs of the REFERENCE's unmodified
vector_store.py (tests/golden/voitta_cases.json, made by tests/golden/make_golden.py)."""
import json
from pathlib import Path

import pytest

from oracle import oracle as O
from golden import make_golden as G
from _parity import assert_same_ranking

GOLD = json.loads((Path(__file__).parent / "golden" / "voitta_cases.json").read_text())


@pytest.fixture(scope="module")
def replayed():
    corpus, queries = G.build_inputs()
    store = O.OracleVectorStore(GOLD["dim"])
    store.get_collection_info = lambda: {"points_count": store.coll.count(None)}
    outs = G.replay(store, GOLD["ops"], corpus, queries, O.ChunkMetadata, {})
    return outs


def test_scenario_is_the_recorded_one():
    corpus, queries = G.build_inputs()
    assert json.loads(json.dumps(G.scenario(corpus, queries))) == GOLD["ops"]


def test_oracle_matches_reference_outputs(replayed):
    assert len(replayed) == len(GOLD["outs"])
    n_search = 0
    for i, (op, got, want) in enumerate(zip(GOLD["ops"], replayed, GOLD["outs"])):
        what = f"op {i} {op}"
        if op["op"] == "search":
            n_search += 1
            # same Python arithmetic on both sides: scores must be bit-equal, ids equal
            # up to permutation inside exact-tie groups (reference order there is hash-seed
            # dependent, vector_store.py:675-689)
            assert_same_ranking(got, want, rel_tol=0.0, what=what)
            for g, w in zip(got, want):
                if g[0] == w[0]:
                    assert g[2:] == w[2:], what
        elif op["op"] in ("get_chunks_by_range", "find_by_source_url"):
            assert json.loads(json.dumps(got)) == want, what
        else:
            assert json.loads(json.dumps(got)) == want, what
    assert n_search >= 40


def test_real_qdrant_client_agrees_when_installed():
    """The day qdrant-client is importable (a driver-provided baseline/_ref, or the package in the image) the golden
    scenario is replayed through the reference over the REAL client and must match the committed outputs — that flips
    the oracle's qdrant half from "parity unpinned" to pinned.  Not installed in this image: skipped, and said so."""
    import pytest
    from golden import make_golden as G
    if not G.real_qdrant_available():
        pytest.skip("qdrant-client is not installed (no network): the qdrant half of the oracle stays pinned only by the "
                    "hand-derived vectors in tests/golden/known_answers.json")
    assert G.check_against_real_qdrant() > 40
