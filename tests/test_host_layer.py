"""Host layer of the drop-in (voitta_rag_b200.vector_store) on a CPU box: the reference's
recorded behaviour (tests/golden) replayed through VectorStoreService over a test-double index.
Checks the 22-method surface, payload/id bookkeeping, filter folding and error conventions."""
import inspect
import json
from pathlib import Path

import pytest

from golden import make_golden as G
from _fake_index import FakeIndex
from _parity import assert_same_ranking
from voitta_rag_b200 import vector_store as VS

GOLD = json.loads((Path(__file__).parent / "golden" / "voitta_cases.json").read_text())


@pytest.fixture()
def store(monkeypatch):
    monkeypatch.setenv("QDRANT_COLLECTION", "host_layer_test")
    monkeypatch.setenv("EMBEDDING_DIMENSION", str(GOLD["dim"]))
    VS._drop_collection("host_layer_test")
    s = VS.VectorStoreService(_index_factory=lambda: FakeIndex(GOLD["dim"]))
    yield s
    VS._drop_collection("host_layer_test")


def check_outputs(outs):
    assert len(outs) == len(GOLD["outs"])
    for i, (op, got, want) in enumerate(zip(GOLD["ops"], outs, GOLD["outs"])):
        what = f"op {i} {op}"
        if op["op"] == "search":
            assert_same_ranking(got, want, rel_tol=1e-5, abs_tol=2e-6, what=what)
            for g, w in zip(got, want):
                if g[0] == w[0]:
                    assert g[2:] == w[2:], what
        else:
            assert json.loads(json.dumps(got)) == want, what


def test_golden_replay_through_host_layer(store):
    corpus, queries = G.build_inputs()
    outs = G.replay(store, GOLD["ops"], corpus, queries, VS.ChunkMetadata, {})
    check_outputs(outs)


def test_surface_matches_reference_signatures():
    """Names, positional order and defaults of SURVEY.md §8(b)."""
    sig = inspect.signature(VS.VectorStoreService.search)
    # the reference's parameters, in order; the one additive parameter comes after them and defaults to None
    assert list(sig.parameters) == ["self", "query_embedding", "limit", "folder_filter", "include_folders",
                                    "exclude_folders", "exclude_index_folders", "sparse_query", "sparse_weight",
                                    "date_start", "date_end", "date_field", "scope_key"]
    assert sig.parameters["scope_key"].default is None
    assert sig.parameters["limit"].default == 10 and sig.parameters["sparse_weight"].default == 0.1
    sig = inspect.signature(VS.VectorStoreService.store_chunks)
    assert list(sig.parameters) == ["self", "chunks", "sparse_vectors", "batch_size"]
    assert sig.parameters["batch_size"].default == 100
    for name in ["find_by_source_url", "set_file_acl", "store_chunks", "delete_by_file", "delete_by_folder",
                 "delete_by_index_folder", "get_file_paths_by_index_folder", "_build_filter", "search",
                 "get_collection_info", "count_by_file", "count_chunks_for_files", "count_chunks_for_folder",
                 "get_folder_stats_batch", "get_stored_page_count", "get_chunks_by_range", "get_file_chunk_counts"]:
        assert callable(getattr(VS.VectorStoreService, name)), name
    fields = [f for f in VS.ChunkMetadata.__dataclass_fields__]
    assert fields == ["file_path", "folder_path", "index_folder", "file_name", "chunk_index", "total_chunks",
                      "start_char", "end_char", "indexed_at", "start_page", "end_page", "source_page_count",
                      "source_created_at", "source_modified_at", "allowed_users", "source_url"]
    assert [f for f in VS.StoredChunk.__dataclass_fields__] == ["id", "text", "metadata", "score"]
    assert VS.get_vector_store() is VS.get_vector_store()


def test_second_instance_sees_same_collection(store):
    corpus, queries = G.build_inputs()
    G.replay(store, GOLD["ops"][:1], corpus, queries, VS.ChunkMetadata, {})
    other = VS.VectorStoreService(_index_factory=lambda: FakeIndex(GOLD["dim"]))   # folders.py:139-141 does this
    assert other.count_by_file(corpus["metas"][0]["file_path"]) == 3
    assert other.delete_by_index_folder("root0") > 0
    assert store.count_by_file(corpus["metas"][0]["file_path"]) in (0, 3)


def test_error_conventions(store):
    assert store.store_chunks([]) == []
    assert store.search([0.0] * GOLD["dim"]) == []              # empty collection
    assert store.count_by_file("x") == 0 and store.count_chunks_for_files([]) == {}
    assert store.get_folder_stats_batch([]) == {} and store.get_stored_page_count("x") is None
    assert store.delete_by_file("x") == 0 and store.find_by_source_url("x") == []
    m = VS.ChunkMetadata("a/f", "a", "a", "f", 0, 1, 0, 1, "t")
    with pytest.raises(ValueError):
        store.store_chunks([("t", [0.0] * 3, m)])               # wrong dimension propagates
    with pytest.raises(ValueError):
        store.store_chunks([("t", [1.0] * GOLD["dim"], m)], [([5, 5], [1.0, 1.0])])   # duplicate sparse index
    store.store_chunks([("t", [1.0] * GOLD["dim"], m)], [([9, 5], [1.0, 2.0])])
    with pytest.raises(ValueError):
        store.search([0.0] * 3)
    assert store._build_filter() is None and store._build_filter(include_folders=[]) is None
    f = store._build_filter(include_folders=["a"], date_end=5, date_field="created")
    assert f.ts_field == 1 and f.ts_hi == 5 and f.scope_bits[0] == 1


def test_snapshot_roundtrip_restores_ids_payloads_and_results(store, tmp_path):
    """save_snapshot / load_snapshot (stand-in for Qdrant's storage volume): after ingest + deletes the
    restored collection answers every helper and every search exactly as the original."""
    corpus, queries = G.build_inputs()
    split = next(i for i, op in enumerate(GOLD["ops"]) if op["op"] == "search")
    while GOLD["ops"][split]["op"] == "search":
        split += 1
    ops = GOLD["ops"]
    ctx = {}
    head = G.replay(store, ops[:split], corpus, queries, VS.ChunkMetadata, ctx)
    info = store.save_snapshot(str(tmp_path))
    assert info["live"] == store.get_collection_info()["points_count"] and Path(info["device_file"]).exists()
    before = {fp: store.count_by_file(fp) for fp in {m["file_path"] for m in corpus["metas"]}}
    searches = [op for op in ops[:split] if op["op"] == "search"]
    want = G.replay(store, searches, corpus, queries, VS.ChunkMetadata, dict(ctx))
    # a different process: forget the collection, restore it from the snapshot
    VS._drop_collection("host_layer_test")
    fresh = VS.VectorStoreService(_index_factory=lambda: FakeIndex(GOLD["dim"]))
    assert fresh.get_collection_info()["points_count"] == 0
    assert fresh.load_snapshot(str(tmp_path), _index_loader=FakeIndex.load) == info["live"]
    assert {fp: fresh.count_by_file(fp) for fp in before} == before
    got = G.replay(fresh, searches, corpus, queries, VS.ChunkMetadata, dict(ctx))
    assert json.loads(json.dumps(got)) == json.loads(json.dumps(want))
    # and it keeps working as a live collection: ingest and delete after the restore
    m = VS.ChunkMetadata("z/new.txt", "z", "z", "new.txt", 0, 1, 0, 1, "t")
    ids = fresh.store_chunks([("t", [1.0] * GOLD["dim"], m)], [([3, 9], [1.0, 2.0])])
    assert len(ids) == 1 and fresh.count_by_file("z/new.txt") == 1 and fresh.delete_by_file("z/new.txt") == 1


def test_concurrent_readers_and_writer(store):
    """The reference's singleton is hit from several threads (MCP tool calls, the indexing worker, the file
    watcher): concurrent store / search / delete / count through the host layer neither raise nor
    corrupt the id / payload bookkeeping."""
    import threading
    dim = GOLD["dim"]
    errors = []

    def meta(i, f):
        return VS.ChunkMetadata(f"root/{f}", "root", "root", f, i, 4, 0, 1, "t")

    def writer(k):
        try:
            for it in range(12):
                f = f"w{k}_{it}.txt"
                ids = store.store_chunks([(f"text {i}", [float((i + k + it) % 7 + 1)] * dim, meta(i, f)) for i in range(4)],
                                         [([3 + i, 11 + k], [1.0, 2.0]) for i in range(4)])
                assert len(ids) == 4 and store.count_by_file(f"root/{f}") == 4
                if it % 3 == 2:
                    assert store.delete_by_file(f"root/{f}") == 4
        except Exception as e:                                  # pragma: no cover - reported below
            errors.append(repr(e))

    def reader():
        try:
            for _ in range(40):
                res = store.search([1.0] * dim, limit=5, sparse_query=([3, 11], [1.0, 1.0]))
                assert len(res) <= 5 and all(r.metadata.folder_path == "root" for r in res)
                store.get_collection_info(); store.get_file_chunk_counts("root")
        except Exception as e:                                  # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=writer, args=(k,)) for k in range(3)] + [threading.Thread(target=reader) for _ in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not errors, errors
    # 3 writers x 12 files x 4 chunks, every third file deleted again
    assert store.get_collection_info()["points_count"] == 3 * (12 - 4) * 4


def test_concurrent_single_queries_are_coalesced_and_identical_to_serial(store):
    """voitta issues B = 1 from several threads (mcp_server.py:474-485).  The submit queue turns callers that overlap
    into one device batch, each with its OWN filter; every caller must get exactly what a serial call returns."""
    import threading
    corpus, queries = G.build_inputs()
    G.replay(store, GOLD["ops"][:2], corpus, queries, VS.ChunkMetadata, {})
    folders = sorted({m["folder_path"] for m in corpus["metas"]})
    calls = []
    for i, (q, sq) in enumerate(queries):
        kw = {}
        if i % 3 == 1:
            kw = {"include_folders": folders[: 1 + i % 4]}
        elif i % 3 == 2:
            kw = {"exclude_folders": folders[:2], "date_start": 1500000000}
        calls.append((list(map(float, q)), (list(sq[0]), list(sq[1])) if i % 4 else None, 5 + (i % 2) * 5, kw))
    serial = [[(c.id, c.score) for c in store.search(q, limit=lim, sparse_query=sq, **kw)] for q, sq, lim, kw in calls]
    assert any(serial)
    before = store.coalescing_stats()
    got = [None] * len(calls)
    errors = []
    gate = threading.Barrier(len(calls))

    def run(i):
        try:
            q, sq, lim, kw = calls[i]
            gate.wait()
            for _ in range(5):
                got[i] = [(c.id, c.score) for c in store.search(q, limit=lim, sparse_query=sq, **kw)]
        except Exception as e:          # pragma: no cover
            errors.append(e)

    th = [threading.Thread(target=run, args=(i,)) for i in range(len(calls))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors
    assert got == serial
    after = store.coalescing_stats()
    assert after["requests"] - before["requests"] == 5 * len(calls)
    assert after["batches"] - before["batches"] <= 5 * len(calls)


def test_scope_key_caches_the_folded_bitset_until_a_new_scope_appears(store):
    corpus, queries = G.build_inputs()
    G.replay(store, GOLD["ops"][:1], corpus, queries, VS.ChunkMetadata, {})
    folders = sorted({m["folder_path"] for m in corpus["metas"]})
    f1 = store._build_filter(include_folders=folders[:3], scope_key=("u", "p", 1))
    f2 = store._build_filter(include_folders=["ignored: the key says the settings did not change"], scope_key=("u", "p", 1))
    assert f2.scope_bits is f1.scope_bits
    f3 = store._build_filter(include_folders=folders[:3], scope_key=("u", "p", 2))      # settings version bumped by the caller
    assert f3.scope_bits is not f1.scope_bits and (f3.scope_bits == f1.scope_bits).all()
    m = VS.ChunkMetadata("brand/new/file.txt", "brand/new", "brand", "file.txt", 0, 1, 0, 1, "t")
    store.store_chunks([("t", [1.0] * GOLD["dim"], m)], [([3], [1.0])])                 # a new scope id: cached bitsets die
    f4 = store._build_filter(include_folders=folders[:3] + ["brand/new"], scope_key=("u", "p", 2))
    assert f4.scope_bits is not f3.scope_bits
    hits = store.search([1.0] * GOLD["dim"], limit=50, include_folders=folders[:3] + ["brand/new"], scope_key=("u", "p", 2))
    assert any(c.metadata.folder_path == "brand/new" for c in hits)


def test_flat_sparse_batches_and_global_idf_weights_are_bit_equal_to_the_per_term_form():
    """engine.flatten_sparse / FlatSparse (one pass over a batch of the reference's (indices, values) pairs) and the
    sharded layer's vectorised IDF weighting: same CSR, same doubles as the per-term formula
    ln((N - df + 0.5) / (df + 0.5) + 1) * value (qdrant local mode's IDF modifier, global statistics)."""
    import math
    import numpy as np
    from voitta_rag_b200 import engine
    from voitta_rag_b200.sharded import ShardedIndex
    rng = np.random.RandomState(3)
    vocab = np.unique(rng.randint(1, 2**32 - 1, size=500, dtype=np.int64))
    sparse = []
    for i in range(97):
        k = int(rng.randint(0, 9))
        if i % 11 == 0:
            sparse.append(None)
        elif i % 5 == 0:                                      # numpy inputs and Python lists mix freely
            sparse.append((rng.choice(vocab, size=k, replace=False), rng.rand(k)))
        else:
            sparse.append(([int(t) for t in rng.choice(vocab, size=k, replace=False)], [float(v) for v in rng.rand(k)]))
    flat = engine.flatten_sparse(sparse, len(sparse))
    assert len(flat) == len(sparse) and flat.indptr[0] == 0
    for i, s in enumerate(sparse):
        t, w = flat[i]
        assert list(t) == ([] if s is None else [int(x) for x in s[0]])
        assert list(w) == ([] if s is None else [float(x) for x in s[1]])
    assert engine.flatten_sparse(flat, len(sparse)) is flat
    with pytest.raises(ValueError):
        engine.flatten_sparse(sparse, len(sparse) + 1)
    with pytest.raises(ValueError):
        engine.flatten_sparse([([1, 2], [1.0])], 1)
    with pytest.raises(ValueError):
        engine.flatten_sparse([([-1], [1.0])], 1)
    sh = ShardedIndex.__new__(ShardedIndex)                   # only the IDF state: no process group, no device
    sh.terms_g = vocab[::2].astype(np.uint32)                 # half of the vocabulary is known to the shards
    sh.df_g = rng.randint(1, 5000, size=len(sh.terms_g)).astype(np.int64)
    sh.n_live_g = 123_457
    sh._idf_cache = {}
    got = sh.idf_weights(sparse)
    df_of = dict(zip(sh.terms_g.tolist(), sh.df_g.tolist()))
    for i, s in enumerate(sparse):
        t, w = got[i]
        if s is None:
            assert len(t) == 0
            continue
        want = [float(v) * math.log((sh.n_live_g - df_of.get(int(x), 0) + 0.5) / (df_of.get(int(x), 0) + 0.5) + 1.0) for x, v in zip(s[0], s[1])]
        assert list(w) == want, f"query {i}"
    # the packed batch built from the flat form equals the one built from the pairs
    q = rng.randn(len(sparse), 8).astype(np.float32)
    a = engine._Packed(8, q, sparse, None, None, 5, 15, 1, 0.1, True)
    b = engine._Packed(8, q, flat, None, None, 5, 15, 1, 0.1, True)
    assert (a.indptr == b.indptr).all() and (a.terms == b.terms).all() and (a.weights == b.weights).all()


def test_qdrant_scroll_facade_serves_the_admin_scan(store):
    """scripts/sync_qdrant_stats.py:29-81 scrolls every point (limit 1000, four payload keys, no vectors) and counts
    chunks per file.  The same loop over ``store.qdrant_facade()`` must see every live chunk exactly once, with the
    payload keys it asks for, across page boundaries and deletions."""
    corpus, queries = G.build_inputs()
    n = len(corpus["texts"])
    chunks = [(corpus["texts"][r], corpus["dense"][r].astype(float).tolist(), VS.ChunkMetadata(**corpus["metas"][r])) for r in range(n)]
    ids = store.store_chunks(chunks, [corpus["sparse"][r] for r in range(n)])
    gone = corpus["metas"][3]["file_path"]
    store.delete_by_file(gone)
    client = store.qdrant_facade()
    total = client.get_collection(store.collection_name).points_count
    assert total == sum(store.get_file_chunk_counts().values()) and 0 < total < n
    for page in (7, 1000):
        seen, stats, offset, pages = [], {}, None, 0
        while True:
            results, offset = client.scroll(collection_name=store.collection_name, limit=page, offset=offset,
                                            with_payload=["file_path", "folder_path", "index_folder", "indexed_at"], with_vectors=False)
            pages += 1
            for pt in results:
                assert set(pt.payload) <= {"file_path", "folder_path", "index_folder", "indexed_at"} and "file_path" in pt.payload
                seen.append(pt.id)
                st = stats.setdefault(pt.payload["file_path"], {"chunk_count": 0, "folder_path": pt.payload.get("folder_path", "")})
                st["chunk_count"] += 1
            if offset is None:
                break
        assert len(seen) == len(set(seen)) == total
        assert gone not in stats
        assert {fp: s_["chunk_count"] for fp, s_ in stats.items()} == store.get_file_chunk_counts()
        assert set(seen) <= set(ids)
        assert pages >= (total + page - 1) // page
    with pytest.raises(ValueError):
        client.scroll(collection_name="another_collection", limit=5)
