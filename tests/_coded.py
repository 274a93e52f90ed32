"""Integer-coded view of a tests/_data.py corpus: the arrays the C ABI (and the C oracle) take."""
from __future__ import annotations

import numpy as np

TS_MISSING = -(2 ** 63)
TS_MIN = -(2 ** 63) + 1
TS_MAX = 2 ** 63 - 1


def code_corpus(corpus):
    metas = corpus["metas"]
    n = len(metas)
    scopes: dict[tuple[str, str], int] = {}
    scope = np.zeros(n, np.uint32)
    created = np.full(n, TS_MISSING, np.int64)
    modified = np.full(n, TS_MISSING, np.int64)
    for r, m in enumerate(metas):
        key = (m["folder_path"], m["index_folder"])
        scope[r] = scopes.setdefault(key, len(scopes))
        if m["source_created_at"] is not None:
            created[r] = m["source_created_at"]
        if m["source_modified_at"] is not None:
            modified[r] = m["source_modified_at"]
    indptr = np.zeros(n + 1, np.int64)
    terms, vals = [], []
    for r, (idx, val) in enumerate(corpus["sparse"]):
        ix = np.asarray(idx, dtype=np.int64)
        order = np.argsort(ix, kind="stable")
        terms.append(ix[order].astype(np.uint32))
        vals.append(np.asarray(val, dtype=np.float32)[order])
        indptr[r + 1] = indptr[r] + len(ix)
    csr = (indptr, np.concatenate(terms) if terms else np.zeros(0, np.uint32),
           np.concatenate(vals) if vals else np.zeros(0, np.float32))
    return {"dense": corpus["dense"], "csr": csr, "scope": scope, "created": created, "modified": modified,
            "scope_list": list(scopes.keys())}


def scope_bits(scope_list, folder_filter=None, include=None, exclude=None, disabled=None):
    """The host-side folding of _build_filter's string clauses (vector_store.py:476-508)."""
    if not (folder_filter or include or exclude or disabled):
        return None
    inc = set(include) if include else None
    exc, dis = set(exclude or ()), set(disabled or ())
    bits = np.zeros(max(1, (len(scope_list) + 31) // 32), np.uint32)
    for sid, (fp, ifp) in enumerate(scope_list):
        if ((not folder_filter or fp == folder_filter) and (inc is None or fp in inc)
                and fp not in exc and ifp not in dis):
            bits[sid >> 5] |= np.uint32(1 << (sid & 31))
    return bits


def passing_rows(coded, flt, alive=None):
    """Row mask of a filter tuple (bits, ts_field, lo, hi) — numpy restatement for property checks."""
    n = len(coded["scope"])
    ok = np.ones(n, bool) if alive is None else alive.astype(bool).copy()
    if flt is None:
        return ok
    bits, field, lo, hi = flt
    if bits is not None:
        s = coded["scope"]
        ok &= ((bits[s >> 5] >> (s & 31)) & 1).astype(bool)
    if field:
        col = coded["created"] if field == 1 else coded["modified"]
        ok &= (col != TS_MISSING) & (col >= lo) & (col <= hi)
    return ok
