"""ctypes wrapper of oracle/liboracle_c.so (the C restatement; TEST INFRASTRUCTURE ONLY —
see oracle_c.c).  Same call shape as voitta_rag_b200.engine.Index.search_batch so parity tests
and bench.py's cpu_baseline feed both sides identical arrays."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle_c.so"
TS_MISSING = -(2 ** 63)


def build(force: bool = False) -> Path:
    if force or not LIB.exists() or LIB.stat().st_mtime < (HERE / "oracle_c.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-B", "liboracle_c.so"], check=True, capture_output=True)
    return LIB


class _Filter(C.Structure):
    _fields_ = [("scope_bits", C.POINTER(C.c_uint32)), ("scope_words", C.c_uint32),
                ("ts_field", C.c_int32), ("ts_lo", C.c_int64), ("ts_hi", C.c_int64)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(str(LIB))
        lib.orc_build.restype = C.c_void_p
        lib.orc_build.argtypes = [C.c_uint64, C.c_int32] + [C.c_void_p] * 8
        lib.orc_free.argtypes = [C.c_void_p]
        lib.orc_df.restype = C.c_uint64
        lib.orc_df.argtypes = [C.c_void_p, C.c_uint32]
        lib.orc_search.argtypes = ([C.c_void_p, C.c_uint32] + [C.c_void_p] * 4 + [C.c_uint32, C.c_void_p, C.c_void_p,
                                   C.c_uint32, C.c_uint32, C.c_int32, C.c_double] + [C.c_void_p] * 9 + [C.c_int32, C.c_int32])
        lib.orc_num_threads.restype = C.c_int
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(_load().orc_num_threads())


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class CorpusC:
    """Integer-coded corpus for the C oracle.  Arrays are kept alive by this object."""

    def __init__(self, dense, sparse_csr=None, scope=None, created=None, modified=None, alive=None):
        self.dense = np.ascontiguousarray(dense, dtype=np.float32)
        self.n, self.dim = self.dense.shape
        self.indptr = self.terms = self.vals = None
        if sparse_csr is not None:
            self.indptr = np.ascontiguousarray(sparse_csr[0], dtype=np.int64)
            self.terms = np.ascontiguousarray(sparse_csr[1], dtype=np.uint32)
            self.vals = np.ascontiguousarray(sparse_csr[2], dtype=np.float32)
        self.scope = None if scope is None else np.ascontiguousarray(scope, dtype=np.uint32)
        self.created = None if created is None else np.ascontiguousarray(created, dtype=np.int64)
        self.modified = None if modified is None else np.ascontiguousarray(modified, dtype=np.int64)
        self.alive = None if alive is None else np.ascontiguousarray(alive, dtype=np.uint8)
        self._h = _load().orc_build(self.n, self.dim, _p(self.dense), _p(self.indptr), _p(self.terms), _p(self.vals),
                                    _p(self.scope), _p(self.created), _p(self.modified), _p(self.alive))

    def __del__(self):
        try:
            if self._h:
                _load().orc_free(self._h)
                self._h = None
        except Exception:
            pass

    def df(self, term: int) -> int:
        return int(_load().orc_df(self._h, int(term)))

    def search_batch(self, queries, sparse=None, filters=None, filter_of=None, limit=10, kprime=None,
                     fusion=1, sparse_weight=0.1, n_threads=0, apply_idf=True):
        """filters: list of (scope_bits|None, ts_field, ts_lo, ts_hi).  fusion: 0 dense, 1 weighted, 2 rrf.
        Returns dict with rows/scores/counts and the branch lists."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        B = q.shape[0]
        if kprime is None:
            kprime = limit * 3 if (sparse is not None and fusion != 0) else limit
        indptr = terms = vals = None
        if sparse is not None:
            lens = [0 if s is None else len(s[0]) for s in sparse]
            indptr = np.zeros(B + 1, np.int64)
            np.cumsum(lens, out=indptr[1:])
            terms = np.zeros(max(1, int(indptr[-1])), np.uint32)
            vals = np.zeros(max(1, int(indptr[-1])), np.float64)
            for i, s in enumerate(sparse):
                if s is not None and lens[i]:
                    terms[indptr[i]:indptr[i + 1]] = np.asarray(s[0], dtype=np.int64).astype(np.uint32)
                    vals[indptr[i]:indptr[i + 1]] = np.asarray(s[1], dtype=np.float64)
        flist = list(filters or [])
        keep = []
        cf = (_Filter * max(1, len(flist)))()
        for i, f in enumerate(flist):
            bits = None if f[0] is None else np.ascontiguousarray(f[0], dtype=np.uint32)
            keep.append(bits)
            cf[i] = _Filter(None if bits is None else bits.ctypes.data_as(C.POINTER(C.c_uint32)),
                            0 if bits is None else bits.size, int(f[1]), int(f[2]), int(f[3]))
        fo = None if filter_of is None else np.ascontiguousarray(filter_of, dtype=np.int32)
        out = {
            "rows": np.zeros((B, limit), np.uint64), "scores": np.zeros((B, limit), np.float64),
            "counts": np.zeros(B, np.int32),
            "dense_rows": np.zeros((B, kprime), np.uint64), "dense_scores": np.zeros((B, kprime), np.float32),
            "dense_counts": np.zeros(B, np.int32),
            "sparse_rows": np.zeros((B, kprime), np.uint64), "sparse_scores": np.zeros((B, kprime), np.float32),
            "sparse_counts": np.zeros(B, np.int32),
        }
        rc = _load().orc_search(self._h, B, _p(q), _p(indptr), _p(terms), _p(vals), len(flist),
                                C.cast(cf, C.c_void_p), _p(fo), limit, kprime, int(fusion), float(sparse_weight),
                                *[_p(out[k]) for k in ("rows", "scores", "counts", "dense_rows", "dense_scores",
                                                       "dense_counts", "sparse_rows", "sparse_scores", "sparse_counts")],
                                int(n_threads), int(bool(apply_idf)))
        if rc != 0:
            raise RuntimeError("orc_search failed")
        return out
