"""Second, independently structured CPU implementation of the hybrid filtered query (BASELINE.md §3 "ref-fast").
TEST INFRASTRUCTURE ONLY — same rule as oracle.py: only tests/, smoke() and bench.py's CPU legs may import it.

oracle.py restates qdrant-client's local mode the way that package computes it (Python loops: per-row payload
filter, two-pointer sparse dot, full argsort); oracle_c.c is a C port of those loops.  This module reaches the
same answers by a different route — integer-coded filter masks, one BLAS product for the dense branch, a
``scipy.sparse`` CSR x vector product for the sparse branch, vectorised fusion — so that an error in the shared
understanding of a formula would have to be made twice, in two shapes, to go unnoticed.  It does NOT pin the
qdrant half of the oracle (qdrant-client itself is still absent, oracle.py header): it cross-checks it.

What it follows (file:line in /root/reference/src/voitta/services/vector_store.py, and the qdrant-client local-mode
modules named in oracle.py's header):
  * filter: must MatchAny(folder_path) / must_not MatchAny / Range on a timestamp key, missing key fails a must range   :462-530
  * dense branch: cosine = dot of the L2-normalised fp32 row and query (norm 0 -> EPSILON)                              :612-617, :640-645
  * sparse branch: IDF-weighted dot, IDF = ln((N - df + 0.5) / (df + 0.5) + 1) over ALL live points, rows without a
    shared index are not results; products and sums in float64 in ascending index order, result rounded to fp32         :647-656
  * k' = 3 * limit per branch, min-max weighted fusion (spread 0 -> 1.0, absent side 0.0) or RRF 1/(2 + rank)           :634-697
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps

EPSILON = np.float32(1.1920929e-7)
TS_MISSING = np.iinfo(np.int64).min


class CorpusFast:
    """Same constructor and ``search_batch`` contract as oracle_c.CorpusC (integer-coded corpus in, dict of padded
    arrays out), so the two can be compared key by key."""

    def __init__(self, dense, sparse_csr=None, scope=None, created=None, modified=None, alive=None):
        self.dense = np.ascontiguousarray(dense, dtype=np.float32)
        self.n, self.dim = self.dense.shape
        norm = np.linalg.norm(self.dense, axis=1).astype(np.float32)
        self.unit = self.dense / np.where(norm != 0, norm, EPSILON)[:, None]
        self.alive = np.ones(self.n, bool) if alive is None else np.asarray(alive).astype(bool)
        self.scope = None if scope is None else np.asarray(scope, dtype=np.int64)
        self.ts = {1: None if created is None else np.asarray(created, dtype=np.int64),
                   2: None if modified is None else np.asarray(modified, dtype=np.int64)}
        self.A = None
        if sparse_csr is not None:
            indptr, terms, vals = (np.asarray(x) for x in sparse_csr)
            self.vocab, col = np.unique(terms.astype(np.int64), return_inverse=True)      # ascending: column order = index order
            self.A = sps.csr_matrix((vals.astype(np.float64), col.astype(np.int64), indptr.astype(np.int64)),
                                    shape=(self.n, len(self.vocab)))
            self.A.sort_indices()
            self.has_vec = np.diff(indptr) > 0
            # document frequency over live points: a row counts for a term when it STORES the index (whatever the value)
            ones = sps.csr_matrix((np.ones(self.A.nnz), self.A.indices, self.A.indptr), shape=self.A.shape)
            self.B = ones
            self.df = np.asarray((sps.diags(self.alive.astype(np.float64)) @ ones).sum(axis=0)).ravel()

    # ---- filter -> boolean mask -------------------------------------------------------------------------------------
    def mask(self, flt) -> np.ndarray:
        m = self.alive.copy()
        if flt is None:
            return m
        bits, field, lo, hi = flt
        if bits is not None:
            if self.scope is None:
                raise ValueError("scope filter on a corpus without scope ids")
            b = np.asarray(bits, dtype=np.uint32)
            word = self.scope >> 5
            ok = word < len(b)
            m &= ok
            m[ok] &= ((b[word[ok]] >> (self.scope[ok] & 31).astype(np.uint32)) & 1).astype(bool)
        if field:
            col = self.ts[int(field)]
            if col is None:
                m[:] = False                                   # the key is missing everywhere: a must range fails
            else:
                m &= (col != TS_MISSING) & (col >= lo) & (col <= hi)
        return m

    # ---- branches ---------------------------------------------------------------------------------------------------
    @staticmethod
    def _top(scores: np.ndarray, ok: np.ndarray, k: int):
        cand = np.flatnonzero(ok)
        if len(cand) == 0:
            return cand, scores[:0]
        s = scores[cand]
        if len(cand) > 4 * k:                                  # cut first, then order exactly (ties: lower row first)
            kth = np.partition(s, len(s) - k)[len(s) - k]
            keep = s >= kth
            cand, s = cand[keep], s[keep]
        order = np.lexsort((cand, -s.astype(np.float64)))[:k]
        return cand[order], s[order]

    def dense_branch(self, Q: np.ndarray, masks, k: int):
        q = np.asarray(Q, dtype=np.float32)
        qn = np.linalg.norm(q, axis=1).astype(np.float32)
        q = q / np.where(qn != 0, qn, EPSILON)[:, None]
        S = self.unit @ q.T                                    # [n, B] fp32, one sgemm
        return [self._top(S[:, i], masks[i], k) for i in range(q.shape[0])]

    def sparse_branch(self, sp, mask, k: int, apply_idf: bool = True):
        if sp is None or len(sp[0]) == 0 or self.A is None:
            return None
        t = np.asarray(sp[0], dtype=np.int64)
        v = np.asarray(sp[1], dtype=np.float64)
        pos = np.searchsorted(self.vocab, t)
        known = (pos < len(self.vocab)) & (self.vocab[np.minimum(pos, len(self.vocab) - 1)] == t)
        w = np.zeros(len(self.vocab))
        ind = np.zeros(len(self.vocab))
        if known.any():
            cols = pos[known]
            if apply_idf:
                # (math.log per term, as the oracles do: np.log's SIMD path may differ from libm in the last bit)
                import math
                N = float(self.alive.sum())
                fac = np.array([math.log((N - self.df[c] + 0.5) / (self.df[c] + 0.5) + 1.0) for c in cols])
                w[cols] = v[known] * fac
            else:
                w[cols] = v[known]
            ind[cols] = 1.0
        scores = (self.A @ w).astype(np.float32)               # csr_matvec: per row, stored (ascending) order, f64 accumulate
        shared = (self.B @ ind) > 0
        return self._top(scores, mask & shared, k)

    # ---- fusion -----------------------------------------------------------------------------------------------------
    @staticmethod
    def _minmax(s: np.ndarray) -> np.ndarray:
        s = s.astype(np.float64)
        spread = s.max() - s.min()
        return (s - s.min()) / spread if spread > 0 else np.ones_like(s)

    def fuse(self, dense, sparse, limit: int, fusion: int, w: float):
        dr, ds = dense
        sr, ss = sparse
        ids = np.concatenate([dr, sr[~np.isin(sr, dr)]])       # first seen: dense list, then sparse-only rows
        if fusion == 2:
            sc = np.zeros(len(ids))
            where = {int(r): i for i, r in enumerate(ids)}
            for lst in (dr, sr):                               # same order of additions as the reference: dense rank first
                for rank, r in enumerate(lst):
                    sc[where[int(r)]] += 1 / (2 + rank)
        else:
            dn = dict(zip(dr.tolist(), self._minmax(ds).tolist())) if len(dr) else {}
            sn = dict(zip(sr.tolist(), self._minmax(ss).tolist())) if len(sr) else {}
            sc = np.array([(1.0 - w) * dn.get(int(r), 0.0) + w * sn.get(int(r), 0.0) for r in ids])
        order = np.argsort(-sc, kind="stable")[:limit]
        return ids[order], sc[order]

    # ---- the batch call ---------------------------------------------------------------------------------------------
    def search_batch(self, queries, sparse=None, filters=None, filter_of=None, limit=10, kprime=None,
                     fusion=1, sparse_weight=0.1, n_threads=0, apply_idf=True):
        Q = np.asarray(queries, dtype=np.float32)
        if Q.ndim == 1:
            Q = Q[None, :]
        B = Q.shape[0]
        if kprime is None:
            kprime = limit * 3 if (sparse is not None and fusion != 0) else limit
        fmasks = [self.mask(f) for f in (filters or [])]
        none = self.mask(None)
        masks = [none if (filter_of is None or filter_of[i] < 0) else fmasks[int(filter_of[i])] for i in range(B)]
        out = {"rows": np.zeros((B, limit), np.uint64), "scores": np.zeros((B, limit), np.float64), "counts": np.zeros(B, np.int32),
               "dense_rows": np.zeros((B, kprime), np.uint64), "dense_scores": np.zeros((B, kprime), np.float32),
               "dense_counts": np.zeros(B, np.int32),
               "sparse_rows": np.zeros((B, kprime), np.uint64), "sparse_scores": np.zeros((B, kprime), np.float32),
               "sparse_counts": np.zeros(B, np.int32)}
        dense = self.dense_branch(Q, masks, kprime)
        for i in range(B):
            dr, ds = dense[i]
            out["dense_rows"][i, :len(dr)], out["dense_scores"][i, :len(dr)], out["dense_counts"][i] = dr, ds, len(dr)
            sb = self.sparse_branch(sparse[i], masks[i], kprime, apply_idf) if (sparse is not None and fusion != 0) else None
            if sb is None:                                     # dense-only query (vector_store.py:612-617): `limit` results
                fr, fs = dr[:limit], ds[:limit].astype(np.float64)
            else:
                sr, ss = sb
                out["sparse_rows"][i, :len(sr)], out["sparse_scores"][i, :len(sr)], out["sparse_counts"][i] = sr, ss, len(sr)
                fr, fs = self.fuse((dr, ds), (sr, ss), limit, fusion, sparse_weight)
            out["rows"][i, :len(fr)], out["scores"][i, :len(fr)], out["counts"][i] = fr, fs, len(fr)
        return out
