"""Test double for voitta_rag_b200.engine.Index backed by the C oracle, so the host layer
(VectorStoreService: payload maps, filter folding, id/row bookkeeping) can be exercised on a
CPU-only box.  Lives in tests/ — the product never imports it."""
from __future__ import annotations

import numpy as np

from oracle import oracle_c
from voitta_rag_b200 import engine


class FakeIndex:
    row_base = 0

    def __init__(self, dim):
        self.dim = dim
        self.dense = np.zeros((0, dim), np.float32)
        self.indptr = np.zeros(1, np.int64)
        self.terms = np.zeros(0, np.uint32)
        self.vals = np.zeros(0, np.float32)
        self.scope = np.zeros(0, np.uint32)
        self.created = np.zeros(0, np.int64)
        self.modified = np.zeros(0, np.int64)
        self.alive = np.zeros(0, np.uint8)
        self._cc = None

    def close(self):
        self._cc = None

    def stats(self):
        return {"n_rows": len(self.dense), "n_live": int(self.alive.sum()), "dim": self.dim, "row_base": 0}

    def save(self, path):
        with open(path, "wb") as f:
            np.savez(f, dim=self.dim, dense=self.dense, indptr=self.indptr, terms=self.terms, vals=self.vals,
                     scope=self.scope, created=self.created, modified=self.modified, alive=self.alive)

    @classmethod
    def load(cls, path):
        z = np.load(path)
        self = cls(int(z["dim"]))
        for k in ("dense", "indptr", "terms", "vals", "scope", "created", "modified", "alive"):
            setattr(self, k, z[k])
        return self

    def upsert(self, dense, sparse_csr=None, scope_id=None, created=None, modified=None):
        n = len(dense)
        first = len(self.dense)
        self.dense = np.vstack([self.dense, np.asarray(dense, np.float32)])
        if sparse_csr is None:
            sparse_csr = (np.zeros(n + 1, np.int64), np.zeros(0, np.uint32), np.zeros(0, np.float32))
        ip = np.asarray(sparse_csr[0], np.int64)
        self.indptr = np.concatenate([self.indptr, self.indptr[-1] + ip[1:] - ip[0]])
        self.terms = np.concatenate([self.terms, np.asarray(sparse_csr[1], np.uint32)])
        self.vals = np.concatenate([self.vals, np.asarray(sparse_csr[2], np.float32)])
        self.scope = np.concatenate([self.scope, np.zeros(n, np.uint32) if scope_id is None else scope_id])
        miss = np.full(n, engine.TS_MISSING, np.int64)
        self.created = np.concatenate([self.created, miss if created is None else created])
        self.modified = np.concatenate([self.modified, miss if modified is None else modified])
        self.alive = np.concatenate([self.alive, np.ones(n, np.uint8)])
        self._cc = None
        return first

    def delete_rows(self, rows):
        self.alive[np.asarray(rows, np.int64)] = 0
        self._cc = None

    def search_batch(self, queries, sparse=None, filters=None, filter_of=None, limit=10, kprime=None,
                     fusion="weighted", sparse_weight=0.1, apply_idf=True, branches=False):
        if self._cc is None:
            self._cc = oracle_c.CorpusC(self.dense, (self.indptr, self.terms, self.vals), self.scope,
                                        self.created, self.modified, self.alive)
        fz = engine.FUSION[fusion] if isinstance(fusion, str) else fusion
        fl = None if not filters else [(f.scope_bits, f.ts_field, f.ts_lo, f.ts_hi) for f in filters]
        out = self._cc.search_batch(queries, sparse, fl, filter_of, limit, kprime, fz, sparse_weight)
        return engine.SearchResult(out["rows"], out["scores"], out["counts"], out["dense_rows"], out["dense_scores"],
                                   out["dense_counts"], out["sparse_rows"], out["sparse_scores"], out["sparse_counts"])
