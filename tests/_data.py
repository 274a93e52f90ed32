"""Seeded synthetic data shared by the CPU and GPU test suites (small scale).

Shapes follow SURVEY.md §8(d): hashed 31-bit sparse term ids with BM25-style values
(the vectors scripts/build_sparse_vectors.py writes), folder tree under index_folders,
epoch timestamps with a few missing.  Dense values are rounded to bf16 so the fp32
oracle and the bf16 GPU index see *identical points*.
"""
from __future__ import annotations

import numpy as np


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (exactly what __float2bfloat16_rn does)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return rounded.astype(np.uint32).view(np.float32).reshape(x.shape)


def make_folders(n_index=3, per_index=4):
    """Returns list of (folder_path, index_folder)."""
    out = []
    for i in range(n_index):
        root = f"root{i}"
        out.append((root, root))
        for j in range(per_index - 1):
            sub = f"{root}/sub{j}" if j % 2 == 0 else f"{root}/sub{j - 1}/deep{j}"
            out.append((sub, root))
    return out


def make_sparse_rows(rng: np.random.RandomState, n: int, vocab: int = 400, lo: int = 3, hi: int = 24,
                     zipf_s: float = 1.07):
    """BM25-style document vectors: distinct hashed term ids, values tf*(k+1)/(tf+k*(1-b+b*len/avg))."""
    term_ids = np.unique(rng.randint(1, 2**31 - 1, size=vocab * 2).astype(np.int64))[:vocab]
    rng.shuffle(term_ids)
    p = 1.0 / np.arange(1, vocab + 1) ** zipf_s
    p /= p.sum()
    rows = []
    k, b, avg_len = 1.2, 0.75, 256.0
    for _ in range(n):
        L = int(rng.randint(lo, hi + 1))
        t = rng.choice(vocab, size=L, replace=False, p=p)
        tf = 1 + rng.geometric(0.6, size=L)
        dl = float(tf.sum())
        val = tf * (k + 1) / (tf + k * (1 - b + b * dl / avg_len))
        rows.append((term_ids[t].tolist(), val.astype(np.float32).astype(float).tolist()))
    return term_ids, rows


def make_corpus(seed: int, n: int, dim: int, n_index=3, per_index=4, vocab=400, chunks_per_file=3,
                missing_ts=0.1, clustered=True):
    """Small labelled corpus: dict of parallel lists ready for store_chunks()."""
    rng = np.random.RandomState(seed)
    folders = make_folders(n_index, per_index)
    if clustered:
        cents = rng.randn(16, dim).astype(np.float32)
        dense = cents[rng.randint(0, 16, size=n)] + 0.6 * rng.randn(n, dim).astype(np.float32)
    else:
        dense = rng.randn(n, dim).astype(np.float32)
    dense = bf16_round(dense * rng.uniform(0.5, 2.0, size=(n, 1)).astype(np.float32))
    term_ids, sparse = make_sparse_rows(rng, n, vocab=vocab)
    fw = 1.0 / np.arange(1, len(folders) + 1)
    fw /= fw.sum()
    metas = []
    t0, t1 = 1420070400, 1767225600
    for r in range(n):
        f = r // chunks_per_file
        fi = int(rng.choice(len(folders), p=fw)) if r % chunks_per_file == 0 else metas[-1]["_fi"]
        folder, index_folder = folders[fi]
        if r % chunks_per_file == 0:
            mod = int(rng.randint(t0, t1))
            cre = mod - int(rng.randint(0, 10_000_000))
            has_mod = rng.rand() >= missing_ts
            has_cre = rng.rand() >= missing_ts
            pdf = rng.rand() < 0.2
            url = f"https://example.org/doc/{f}" if rng.rand() < 0.15 else None
        else:
            prev = metas[-1]
            mod, cre, has_mod, has_cre, pdf, url = (prev["_mod"], prev["_cre"], prev["_hm"], prev["_hc"],
                                                    prev["_pdf"], prev["source_url"])
        ci = r % chunks_per_file
        metas.append({
            "file_path": f"{folder}/file{f}.md", "folder_path": folder, "index_folder": index_folder,
            "file_name": f"file{f}.md", "chunk_index": ci, "total_chunks": chunks_per_file,
            "start_char": ci * 500, "end_char": ci * 500 + 512, "indexed_at": "2025-01-01T00:00:00",
            "start_page": (ci + 1) if pdf else None, "end_page": (ci + 2) if pdf else None,
            "source_page_count": 7 if pdf else None,
            "source_created_at": cre if has_cre else None,
            "source_modified_at": mod if has_mod else None,
            "allowed_users": None, "source_url": url,
            "_fi": fi, "_mod": mod, "_cre": cre, "_hm": has_mod, "_hc": has_cre, "_pdf": pdf,
        })
    texts = [f"chunk {r} of {m['file_path']}" for r, m in enumerate(metas)]
    metas = [{k: v for k, v in m.items() if not k.startswith("_")} for m in metas]
    return {"dense": dense, "sparse": sparse, "metas": metas, "texts": texts,
            "folders": folders, "term_ids": term_ids}


def make_queries(seed: int, corpus: dict, nq: int, noise: float = 0.5, nnz=(3, 12), bf16=True):
    """Dense query = a corpus row + noise; sparse query = a few terms of that row + one random term,
    values 1.0 (what fastembed Qdrant/bm25 emits for queries)."""
    rng = np.random.RandomState(seed)
    dense, sparse, term_ids = corpus["dense"], corpus["sparse"], corpus["term_ids"]
    n, dim = dense.shape
    qs = []
    for _ in range(nq):
        r = int(rng.randint(0, n))
        v = dense[r] / max(float(np.linalg.norm(dense[r])), 1e-9)
        q = v + noise * rng.randn(dim).astype(np.float32) / np.sqrt(dim)
        q = (q / np.linalg.norm(q)).astype(np.float32)
        if bf16:
            q = bf16_round(q)
        terms = list(sparse[r][0])
        k = min(len(terms), int(rng.randint(nnz[0], nnz[1] + 1)))
        pick = [terms[i] for i in rng.choice(len(terms), size=k, replace=False)]
        extra = int(term_ids[rng.randint(0, len(term_ids))])
        if extra not in pick:
            pick.append(extra)
        qs.append((q, (pick, [1.0] * len(pick))))
    return qs
