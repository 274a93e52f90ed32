"""CPU ORACLE for the voitta-rag retrieval hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (voitta-rag_b200/) never does.

What is restated here, and from where
-------------------------------------
(1) voitta's own code, restated from the reference tree (authoritative, file:line
    into /root/reference/src/voitta/services/vector_store.py):
      * point construction in ``store_chunks``            :233-317
      * ``_build_filter``                                  :462-530
      * ``search`` dispatch (hybrid iff indices non-empty) :560-619
      * ``_hybrid_search`` min-max weighted fusion         :621-697
      * ``_result_to_chunk`` (index_folder fallback)       :532-558
      * deletes / counts / scroll helpers                  :163-231, 319-460, 699-1016
    This part IS pinned: tests/golden/make_golden.py imports the reference's
    real vector_store.py in the build container (with qdrant_client replaced by
    a stub that delegates to ``LocalCollection`` below), records its outputs in
    tests/golden/*.json, and tests/test_oracle_golden.py checks this restatement
    against those recorded outputs.

(2) The arithmetic voitta delegates to the third-party package ``qdrant-client``
    (constraint ``>=1.7.0`` in /root/reference/pyproject.toml:31, NOT vendored,
    no lock file, not installable here: no network).  ``LocalCollection`` restates
    the published algorithm of qdrant-client's local (":memory:") mode
    (qdrant_client/local/local_collection.py, distances.py, sparse_distances.py,
    payload_filters.py, hybrid/fusion.py) from the call sites
    vector_store.py:89-100 (COSINE + sparse "bm25" with Modifier.IDF), :313
    (upsert), :612-617/:640-656 (query_points), :171ff (scroll), :326ff (count),
    :340ff (delete), :218 (set_payload).
    **PARITY UNPINNED for this part**: the reference has no test or golden vector
    on this path and qdrant-client cannot be run here, so the
    restatement is anchored on the documented semantics only.

Deliberate, documented deviations (both are places where the reference's own
order is arbitrary, so any order is "correct"):
  * exact score ties inside one branch: qdrant uses ``np.argsort`` (introsort,
    unstable); we order ties by insertion row ascending (``stable_ties=True``).
  * exact ties of the fused score: the reference sorts a list built by iterating
    a Python ``set`` of uuid strings (hash-seed dependent, vector_store.py:675-689);
    we iterate in first-seen order (dense list, then sparse-only ids).
"""
from __future__ import annotations

import math
import uuid
from dataclasses import dataclass, field
from typing import Any, Iterable

import numpy as np

EPSILON = 1.1920929e-7  # qdrant_client/local/distances.py: EPSILON
SPARSE_VECTOR_NAME = "bm25"  # sparse_embedding.py:9


# --------------------------------------------------------------------------------------
# Filter AST (stands in for qdrant_client.http.models.{Filter,FieldCondition,...})
# --------------------------------------------------------------------------------------
@dataclass
class MatchValue:
    value: Any


@dataclass
class MatchAny:
    any: list


@dataclass
class Range:
    gte: float | None = None
    lte: float | None = None
    gt: float | None = None
    lt: float | None = None


@dataclass
class FieldCondition:
    key: str
    match: Any = None
    range: Range | None = None


@dataclass
class Filter:
    must: list | None = None
    must_not: list | None = None
    should: list | None = None


@dataclass
class SparseVector:
    indices: list
    values: list


@dataclass
class ScoredPoint:
    id: str
    score: float
    payload: dict
    row: int = -1  # insertion row (oracle-only convenience for parity checks)


@dataclass
class Record:
    id: str
    payload: dict


def _values_at(payload: dict, key: str) -> list:
    """payload_filters.py value_by_key: missing / None -> no values; lists flatten."""
    if key not in payload or payload[key] is None:
        return []
    v = payload[key]
    return list(v) if isinstance(v, (list, tuple)) else [v]


def check_condition(cond: FieldCondition, payload: dict) -> bool:
    """payload_filters.py check_condition / check_match / check_range.

    A missing key yields no values, so every condition on it is False
    (``must`` fails, ``must_not`` passes)."""
    values = _values_at(payload, cond.key)
    if cond.match is not None:
        if isinstance(cond.match, MatchValue):
            return any(v == cond.match.value for v in values)
        if isinstance(cond.match, MatchAny):
            return any(v in cond.match.any for v in values)
        raise TypeError(cond.match)
    if cond.range is not None:
        r = cond.range

        def ok(v):
            if not isinstance(v, (int, float)) or isinstance(v, bool):
                return False
            return ((r.lt is None or v < r.lt) and (r.gt is None or v > r.gt)
                    and (r.lte is None or v <= r.lte) and (r.gte is None or v >= r.gte))

        return any(ok(v) for v in values)
    return False


def check_filter(flt: Filter | None, payload: dict) -> bool:
    """payload_filters.py check_filter: must = all, must_not = none, should = any."""
    if flt is None:
        return True
    if flt.must is not None and not all(check_condition(c, payload) for c in flt.must):
        return False
    if flt.must_not is not None and any(check_condition(c, payload) for c in flt.must_not):
        return False
    if flt.should is not None and len(flt.should) > 0:
        if not any(check_condition(c, payload) for c in flt.should):
            return False
    return True


# --------------------------------------------------------------------------------------
# qdrant-client local mode, restated (PARITY UNPINNED, see module docstring)
# --------------------------------------------------------------------------------------
def sparse_dot_product(q_idx, q_val, d_idx, d_val):
    """sparse_distances.py sparse_dot_product: two-pointer merge over index-sorted
    vectors, Python-float (f64) accumulation, np.float32 result, None if no overlap."""
    result = 0.0
    i = j = 0
    overlap = False
    nq, nd = len(q_idx), len(d_idx)
    while i < nq and j < nd:
        a, b = q_idx[i], d_idx[j]
        if a == b:
            overlap = True
            result += q_val[i] * d_val[j]
            i += 1
            j += 1
        elif a < b:
            i += 1
        else:
            j += 1
    return np.float32(result) if overlap else None


def _sort_sparse(indices, values):
    """local_collection.py sort_sparse_vector: ascending by index."""
    idx = [int(x) for x in indices]
    val = [float(x) for x in values]
    if len(idx) != len(val):
        raise ValueError("sparse indices/values length mismatch")
    if len(set(idx)) != len(idx):
        raise ValueError("sparse vector indices must be unique")
    order = sorted(range(len(idx)), key=idx.__getitem__)
    return [idx[k] for k in order], [val[k] for k in order]


class LocalCollection:
    """Restatement of qdrant_client.local.LocalCollection for ONE collection created
    as vector_store.py:89-100 does: unnamed dense vector (size=dim, COSINE) and a
    sparse vector "bm25" with Modifier.IDF (if ``has_sparse``)."""

    def __init__(self, dim: int, has_sparse: bool = True, stable_ties: bool = True):
        self.dim = dim
        self.has_sparse = has_sparse
        self.stable_ties = stable_ties
        self.ids: dict[str, int] = {}       # id -> row
        self.ids_inv: list[str] = []        # row -> id
        self.payload: list[dict] = []
        self.deleted: list[bool] = []
        self.dense = np.zeros((0, dim), dtype=np.float32)
        self.sparse: list[tuple[list, list] | None] = []
        self.df: dict[int, int] = {}        # sparse_vectors_idf["bm25"]

    # ---- write side -----------------------------------------------------------------
    def _df_add(self, sv, sign):
        if sv is None:
            return
        for t in sv[0]:
            self.df[t] = self.df.get(t, 0) + sign
            if self.df[t] == 0:
                del self.df[t]

    def upsert(self, pid: str, dense, sparse=None, payload=None) -> int:
        """local_collection.py _upsert_point/_add_point/_update_point.
        COSINE: vector is L2-normalised in f32 at insert; norm 0 -> kept as is."""
        v = np.asarray(dense, dtype=np.float32)
        if v.shape != (self.dim,):
            raise ValueError(f"dense vector has shape {v.shape}, expected ({self.dim},)")
        nrm = np.linalg.norm(v)
        v = v / nrm if nrm > 0 else v
        sv = _sort_sparse(*sparse) if (sparse is not None and self.has_sparse) else None
        if pid in self.ids:                      # overwrite in place
            r = self.ids[pid]
            if not self.deleted[r]:
                self._df_add(self.sparse[r], -1)
            self.dense[r] = v
            self.sparse[r] = sv
            self.payload[r] = dict(payload or {})
            self.deleted[r] = False
        else:
            r = len(self.ids_inv)
            self.ids[pid] = r
            self.ids_inv.append(pid)
            if r >= self.dense.shape[0]:
                grow = max(1024, self.dense.shape[0])
                self.dense = np.vstack([self.dense, np.zeros((grow, self.dim), np.float32)])
            self.dense[r] = v
            self.sparse.append(sv)
            self.payload.append(dict(payload or {}))
            self.deleted.append(False)
        self._df_add(sv, +1)
        return r

    def _live_rows(self, flt: Filter | None) -> list[int]:
        return [r for r in range(len(self.ids_inv))
                if not self.deleted[r] and check_filter(flt, self.payload[r])]

    def count(self, flt: Filter | None = None) -> int:
        return len(self._live_rows(flt))

    def delete(self, flt: Filter) -> int:
        rows = self._live_rows(flt)
        for r in rows:
            self.deleted[r] = True
            self._df_add(self.sparse[r], -1)
        return len(rows)

    def set_payload(self, payload: dict, flt: Filter) -> int:
        rows = self._live_rows(flt)
        for r in rows:
            self.payload[r].update(payload)
        return len(rows)

    def scroll(self, flt: Filter | None = None, limit: int = 10, offset: str | None = None):
        """local_collection.py scroll: matching points in ascending id order;
        returns (records, next_offset|None)."""
        rows = sorted(self._live_rows(flt), key=lambda r: self.ids_inv[r])
        if offset is not None:
            rows = [r for r in rows if self.ids_inv[r] >= offset]
        page, rest = rows[:limit], rows[limit:]
        recs = [Record(self.ids_inv[r], self.payload[r]) for r in page]
        return recs, (self.ids_inv[rest[0]] if rest else None)

    # ---- read side ------------------------------------------------------------------
    def _select(self, scores: np.ndarray, mask: np.ndarray, limit: int) -> list[ScoredPoint]:
        """local_collection.py search(): argsort descending, skip masked / -inf, stop at limit."""
        n = len(self.ids_inv)
        if self.stable_ties:
            order = np.lexsort((np.arange(n), -scores.astype(np.float64)))
        else:
            order = np.argsort(scores)[::-1]
        out = []
        for r in order:
            if len(out) >= limit:
                break
            s = scores[r]
            if not mask[r] or s == -np.inf or np.isnan(s):
                continue
            out.append(ScoredPoint(self.ids_inv[r], float(s), self.payload[r], int(r)))
        return out

    def _mask(self, flt):
        n = len(self.ids_inv)
        return np.array([(not self.deleted[r]) and check_filter(flt, self.payload[r])
                         for r in range(n)], dtype=bool)

    def dense_scores(self, query) -> np.ndarray:
        """distances.py cosine_similarity: re-normalise stored rows and the query in f32
        (norm 0 -> divide by EPSILON), then np.dot."""
        n = len(self.ids_inv)
        q = np.asarray(query, dtype=np.float32).copy()
        if q.shape != (self.dim,):
            raise ValueError(f"query has shape {q.shape}, expected ({self.dim},)")
        assert not np.isnan(q).any(), "Query vector must not contain NaN"
        V = self.dense[:n].copy()
        vn = np.linalg.norm(V, axis=-1)[:, np.newaxis]
        V /= np.where(vn != 0.0, vn, EPSILON)
        qn = np.linalg.norm(q)
        q /= np.where(qn != 0.0, qn, EPSILON)
        return np.dot(V, q).astype(np.float32) if n else np.zeros(0, np.float32)

    def query_dense(self, query, limit: int, flt: Filter | None = None) -> list[ScoredPoint]:
        return self._select(self.dense_scores(query), self._mask(flt), limit)

    def idf(self, term: int, n_points: int) -> float:
        """local_collection.py _compute_idf."""
        df = self.df.get(term, 0)
        return math.log((n_points - df + 0.5) / (df + 0.5) + 1.0)

    def rescore_idf(self, indices, values):
        """local_collection.py _rescore_idf: N = all live points (filter independent)."""
        n_points = self.count(None)
        return [float(v) * self.idf(int(t), n_points) for t, v in zip(indices, values)]

    def sparse_scores(self, indices, values) -> np.ndarray:
        """sparse_distances.py calculate_distance_sparse with the IDF modifier applied to
        the query: rows without the sparse vector or without overlap score -inf."""
        q_idx, q_val = _sort_sparse(indices, values)
        q_val = self.rescore_idf(q_idx, q_val)
        n = len(self.ids_inv)
        scores = np.full(n, -np.inf, dtype=np.float32)
        for r in range(n):
            sv = self.sparse[r]
            if sv is None:
                continue
            s = sparse_dot_product(q_idx, q_val, sv[0], sv[1])
            if s is not None:
                scores[r] = s
        return scores

    def query_sparse(self, indices, values, limit: int, flt: Filter | None = None):
        if not self.has_sparse:
            raise ValueError("collection has no sparse vector 'bm25'")
        return self._select(self.sparse_scores(indices, values), self._mask(flt), limit)


def reciprocal_rank_fusion(responses: list[list[ScoredPoint]], limit: int) -> list[tuple[str, float, ScoredPoint]]:
    """qdrant_client/hybrid/fusion.py reciprocal_rank_fusion: score = sum 1/(2+pos);
    Python's stable ``sorted(..., reverse=True)`` keeps first-seen order on ties."""
    scores: dict[str, float] = {}
    pile: dict[str, ScoredPoint] = {}
    for response in responses:
        for i, sp in enumerate(response):
            if sp.id in scores:
                scores[sp.id] += 1 / (2 + i)
            else:
                pile[sp.id] = sp
                scores[sp.id] = 1 / (2 + i)
    ranked = sorted(scores.items(), key=lambda it: it[1], reverse=True)
    return [(pid, sc, pile[pid]) for pid, sc in ranked[:limit]]


# --------------------------------------------------------------------------------------
# voitta's VectorStoreService, restated over LocalCollection (PINNED by tests/golden)
# --------------------------------------------------------------------------------------
@dataclass
class ChunkMetadata:  # vector_store.py:18-41
    file_path: str
    folder_path: str
    index_folder: str
    file_name: str
    chunk_index: int
    total_chunks: int
    start_char: int
    end_char: int
    indexed_at: str
    start_page: int | None = None
    end_page: int | None = None
    source_page_count: int | None = None
    source_created_at: int | None = None
    source_modified_at: int | None = None
    allowed_users: list | None = None
    source_url: str | None = None


@dataclass
class StoredChunk:  # vector_store.py:44-51
    id: str
    text: str
    metadata: ChunkMetadata
    score: float | None = None


def build_payload(text: str, m) -> dict:
    """vector_store.py:259-288 — optional fields only when not None."""
    p = {
        "text": text, "file_path": m.file_path, "folder_path": m.folder_path,
        "index_folder": m.index_folder, "file_name": m.file_name,
        "chunk_index": m.chunk_index, "total_chunks": m.total_chunks,
        "start_char": m.start_char, "end_char": m.end_char, "indexed_at": m.indexed_at,
    }
    for k in ("start_page", "end_page", "source_page_count", "source_created_at",
              "source_modified_at", "allowed_users", "source_url"):
        v = getattr(m, k)
        if v is not None:
            p[k] = v
    return p


def payload_to_chunk(pid: str, payload: dict, score) -> StoredChunk:
    """vector_store.py:532-558 _result_to_chunk (index_folder falls back to folder_path)."""
    return StoredChunk(
        id=str(pid), text=payload["text"],
        metadata=ChunkMetadata(
            file_path=payload["file_path"], folder_path=payload["folder_path"],
            index_folder=payload.get("index_folder", payload["folder_path"]),
            file_name=payload["file_name"], chunk_index=payload["chunk_index"],
            total_chunks=payload["total_chunks"], start_char=payload["start_char"],
            end_char=payload["end_char"], indexed_at=payload["indexed_at"],
            start_page=payload.get("start_page"), end_page=payload.get("end_page"),
            source_page_count=payload.get("source_page_count"),
            source_created_at=payload.get("source_created_at"),
            source_modified_at=payload.get("source_modified_at"),
            allowed_users=payload.get("allowed_users"), source_url=payload.get("source_url"),
        ),
        score=score,
    )


def build_filter(folder_filter=None, include_folders=None, exclude_folders=None,
                 exclude_index_folders=None, date_start=None, date_end=None,
                 date_field=None) -> Filter | None:
    """vector_store.py:462-530 _build_filter."""
    must, must_not = [], []
    if folder_filter:
        must.append(FieldCondition(key="folder_path", match=MatchValue(value=folder_filter)))
    if include_folders:
        must.append(FieldCondition(key="folder_path", match=MatchAny(any=list(include_folders))))
    if exclude_folders:
        for f in exclude_folders:
            must_not.append(FieldCondition(key="folder_path", match=MatchValue(value=f)))
    if exclude_index_folders:
        for f in exclude_index_folders:
            must_not.append(FieldCondition(key="index_folder", match=MatchValue(value=f)))
    if date_start is not None or date_end is not None:
        field_map = {"created": "source_created_at", "modified": "source_modified_at"}
        key = field_map.get(date_field, "source_modified_at") if date_field else "source_modified_at"
        must.append(FieldCondition(key=key, range=Range(gte=date_start, lte=date_end)))
    if must or must_not:
        return Filter(must=must or None, must_not=must_not or None)
    return None


def weighted_fusion(dense: list[ScoredPoint], sparse: list[ScoredPoint], limit: int,
                    sparse_weight: float) -> list[tuple[str, float, ScoredPoint]]:
    """vector_store.py:634, 659-697: per-list min-max normalisation (spread==0 -> 1.0),
    final = (1-w)*d + w*s with the absent side 0.0, stable sort descending, top ``limit``.
    Iteration order over the id union is first-seen (documented deviation)."""
    dense_weight = 1.0 - sparse_weight

    def normalize(results):
        if not results:
            return {}
        scores = [r.score for r in results]
        min_s, max_s = min(scores), max(scores)
        spread = max_s - min_s
        return {str(r.id): ((r.score - min_s) / spread if spread > 0 else 1.0, r) for r in results}

    dn, sn = normalize(dense), normalize(sparse)
    all_ids = list(dn.keys()) + [k for k in sn.keys() if k not in dn]
    combined = []
    for pid in all_ids:
        d = dn[pid][0] if pid in dn else 0.0
        s = sn[pid][0] if pid in sn else 0.0
        combined.append((dense_weight * d + sparse_weight * s, pid,
                         dn[pid][1] if pid in dn else sn[pid][1]))
    combined.sort(key=lambda x: x[0], reverse=True)
    return [(pid, sc, r) for sc, pid, r in combined[:limit]]


class OracleVectorStore:
    """voitta VectorStoreService restated over LocalCollection.  ``fusion`` selects
    voitta's weighted fusion (default; vector_store.py:659-697) or Qdrant's RRF
    (BASELINE.json configs 2-5)."""

    def __init__(self, dim: int, has_sparse: bool = True, fusion: str = "weighted"):
        self.dimension = dim
        self.coll = LocalCollection(dim, has_sparse=has_sparse)
        self._has_sparse = has_sparse
        self.fusion = fusion

    # ---- write ----------------------------------------------------------------------
    def store_chunks(self, chunks, sparse_vectors=None, batch_size: int = 100, ids=None) -> list[str]:
        """vector_store.py:233-317.  ``ids`` (oracle-only) pins the uuid4 ids for tests."""
        if not chunks:
            return []
        out = []
        for idx, (text, embedding, metadata) in enumerate(chunks):
            pid = ids[idx] if ids is not None else str(uuid.uuid4())
            sv = sparse_vectors[idx] if (sparse_vectors and idx < len(sparse_vectors)) else None
            self.coll.upsert(pid, embedding, sv, build_payload(text, metadata))
            out.append(pid)
        return out

    def _by(self, key, value) -> Filter:
        return Filter(must=[FieldCondition(key=key, match=MatchValue(value=value))])

    def delete_by_file(self, file_path):            # :319-355
        return self.coll.delete(self._by("file_path", file_path))

    def delete_by_folder(self, folder_path):        # :357-393
        return self.coll.delete(self._by("folder_path", folder_path))

    def delete_by_index_folder(self, index_folder):  # :395-434
        return self.coll.delete(self._by("index_folder", index_folder))

    def set_file_acl(self, file_path, allowed_users):  # :216-231
        self.coll.set_payload({"allowed_users": allowed_users}, self._by("file_path", file_path))

    def count_by_file(self, file_path):             # :712-728
        return self.coll.count(self._by("file_path", file_path))

    def _scroll_all(self, flt, page):
        offset = None
        while True:
            recs, offset = self.coll.scroll(flt, limit=page, offset=offset)
            yield from recs
            if offset is None:
                break

    def find_by_source_url(self, source_url):       # :163-214
        chunks = [payload_to_chunk(r.id, r.payload, None)
                  for r in self._scroll_all(self._by("source_url", source_url), 100)]
        chunks.sort(key=lambda c: c.metadata.chunk_index)
        return chunks

    def get_file_paths_by_index_folder(self, index_folder):  # :436-460
        return {r.payload["file_path"] for r in self._scroll_all(self._by("index_folder", index_folder), 1000)}

    def get_chunks_by_range(self, file_path, first_chunk, last_chunk):  # :898-977
        chunks = [payload_to_chunk(r.id, r.payload, None)
                  for r in self._scroll_all(self._by("file_path", file_path), 100)
                  if first_chunk <= r.payload["chunk_index"] <= last_chunk]
        chunks.sort(key=lambda c: c.metadata.chunk_index)
        return chunks

    def get_stored_page_count(self, file_path):     # :869-896
        recs, _ = self.coll.scroll(self._by("file_path", file_path), limit=1)
        if recs and recs[0].payload.get("source_page_count"):
            return recs[0].payload["source_page_count"]
        return None

    def count_chunks_for_files(self, file_paths):   # :730-775
        if not file_paths:
            return {}
        flt = Filter(must=[FieldCondition(key="file_path", match=MatchAny(any=list(file_paths)))])
        out: dict[str, int] = {}
        for r in self._scroll_all(flt, 1000):
            fp = r.payload.get("file_path", "")
            out[fp] = out.get(fp, 0) + 1
        return out

    def count_chunks_for_folder(self, folder_path):  # :777-814
        prefix = folder_path + "/" if folder_path else ""
        fc: dict[str, int] = {}
        for r in self._scroll_all(None, 1000):
            fp = r.payload.get("file_path", "")
            if fp.startswith(prefix) or (not prefix and "/" not in fp):
                fc[fp] = fc.get(fp, 0) + 1
        return len(fc), sum(fc.values())

    def get_folder_stats_batch(self, folder_paths):  # :816-867
        if not folder_paths:
            return {}
        prefixes = [(fp, fp + "/" if fp else "") for fp in folder_paths]
        ff: dict[str, dict[str, int]] = {fp: {} for fp in folder_paths}
        for r in self._scroll_all(None, 1000):
            path = r.payload.get("file_path", "")
            for folder, prefix in prefixes:
                if path.startswith(prefix) or (not prefix and "/" not in path):
                    ff[folder][path] = ff[folder].get(path, 0) + 1
        return {fp: (len(f), sum(f.values())) for fp, f in ff.items()}

    def get_file_chunk_counts(self, folder_prefix=""):  # :979-1016
        out: dict[str, int] = {}
        for r in self._scroll_all(None, 1000):
            fp = r.payload.get("file_path", "")
            if folder_prefix and not fp.startswith(folder_prefix):
                continue
            out[fp] = out.get(fp, 0) + 1
        return out

    # ---- read -----------------------------------------------------------------------
    def search_branches(self, query_embedding, limit, flt, sparse_query):
        """The two ``query_points`` calls of vector_store.py:640-656 (k' = 3*limit)."""
        k = limit * 3
        dense = self.coll.query_dense(query_embedding, k, flt)
        sparse = self.coll.query_sparse(sparse_query[0], sparse_query[1], k, flt)
        return dense, sparse

    def search(self, query_embedding, limit=10, folder_filter=None, include_folders=None,
               exclude_folders=None, exclude_index_folders=None, sparse_query=None,
               sparse_weight=0.1, date_start=None, date_end=None, date_field=None):
        """vector_store.py:560-619 (+ :621-697 for the hybrid branch)."""
        flt = build_filter(folder_filter, include_folders, exclude_folders, exclude_index_folders,
                           date_start=date_start, date_end=date_end, date_field=date_field)
        if sparse_query and self._has_sparse:
            indices, values = sparse_query
            if indices:
                dense, sparse = self.search_branches(query_embedding, limit, flt, (indices, values))
                if self.fusion == "rrf":
                    fused = reciprocal_rank_fusion([dense, sparse], limit)
                else:
                    fused = weighted_fusion(dense, sparse, limit, sparse_weight)
                return [payload_to_chunk(pid, r.payload, sc) for pid, sc, r in fused]
        res = self.coll.query_dense(query_embedding, limit, flt)
        return [payload_to_chunk(r.id, r.payload, r.score) for r in res]
