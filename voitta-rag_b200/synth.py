"""Seeded synthetic corpora of BASELINE.json's shapes, generated ON THE DEVICE (torch is used for
RNG and sorting only — plumbing), so that 10M+ row corpora never touch host RAM.  SURVEY.md §8(d):

  dense   "C": 4096 unit centroids + 0.35*N(0,1)/sqrt(d) noise, normalised, rounded to bf16
          "U": iid N(0,1)^d normalised (worst case for threshold pruning)
  sparse  vocabulary 2^20 term ranks mapped by a fixed injection into [0, 2^31) (hashed ids, as
          fastembed's Qdrant/bm25 emits); per row L ~ clip(lognormal(ln 48, 0.5), 4, 256) terms drawn
          Zipf(1.07); tf ~ 1 + Geometric(0.6); value = tf*(k+1)/(tf + k*(1-b+b*len/avg)), k=1.2,
          b=0.75, avg=256 — the document side of build_sparse_vectors.py's BM25 vectors
  scope   4096 scope ids (folder_path x index_folder pairs), rows assigned Zipf(1.0)
  time    modified uniform in [2015-01-01, 2026-01-01), 5 % missing; created = modified - U[0, 1e7]
  queries a corpus row + 0.5*noise (dense); 3..12 of that row's terms + 1 random term, values 1.0
"""
from __future__ import annotations

import numpy as np
import torch

TS_MISSING = -(2 ** 63)
T0, T1 = 1420070400, 1767225600
VOCAB = 1 << 20
N_SCOPES = 4096


def hash_term(rank: torch.Tensor) -> torch.Tensor:
    """Fixed injection of term ranks into [0, 2^31): multiplication by an odd constant mod 2^31."""
    return (rank.to(torch.int64) * 2654435761 + 12345) & 0x7FFFFFFF


def _zipf_cdf(n: int, s: float, device) -> torch.Tensor:
    p = 1.0 / torch.arange(1, n + 1, device=device, dtype=torch.float64) ** s
    return torch.cumsum(p / p.sum(), 0)


def dense_rows(n: int, dim: int, seed: int, device, dist: str = "C", chunk: int = 1 << 20) -> torch.Tensor:
    """[n, dim] bf16 on the device."""
    g = torch.Generator(device=device).manual_seed(1234 + seed)
    out = torch.empty((n, dim), dtype=torch.bfloat16, device=device)
    cents = None
    if dist == "C":
        gc = torch.Generator(device=device).manual_seed(99)
        cents = torch.nn.functional.normalize(torch.randn((4096, dim), generator=gc, device=device), dim=1)
    for r0 in range(0, n, chunk):
        m = min(chunk, n - r0)
        x = torch.randn((m, dim), generator=g, device=device)
        if dist == "C":
            idx = torch.randint(0, 4096, (m,), generator=g, device=device)
            x = cents[idx] + 0.35 * x / dim ** 0.5
        out[r0:r0 + m] = torch.nn.functional.normalize(x, dim=1).to(torch.bfloat16)
    return out


def sparse_rows(n: int, seed: int, device, vocab: int = VOCAB, mean_len: float = 48.0):
    """CSR on the device: indptr int64 [n+1], terms int32 (hashed ids, ascending per row), vals float32."""
    g = torch.Generator(device=device).manual_seed(777 + seed)
    L = torch.exp(torch.randn(n, generator=g, device=device) * 0.5 + np.log(mean_len)).clamp(4, 256).long()
    total = int(L.sum().item())
    row = torch.repeat_interleave(torch.arange(n, device=device), L)
    cdf = _zipf_cdf(vocab, 1.07, device)
    u = torch.rand(total, generator=g, device=device, dtype=torch.float64)
    rank = torch.searchsorted(cdf, u).clamp_(max=vocab - 1)
    del u
    key = (row << 32) | hash_term(rank)
    del rank, row
    key = torch.unique(key, sorted=True)                     # distinct terms per row, ascending id
    row = key >> 32
    terms = (key & 0x7FFFFFFF).to(torch.int32)
    del key
    nnz = terms.numel()
    tf = 1.0 + torch.floor(torch.log(torch.rand(nnz, generator=g, device=device).clamp_min(1e-12)) / np.log(0.4))
    dl = torch.zeros(n, device=device).index_add_(0, row, tf)
    k, b, avg = 1.2, 0.75, 256.0
    vals = (tf * (k + 1.0) / (tf + k * (1.0 - b + b * dl[row] / avg))).to(torch.float32)
    counts = torch.bincount(row, minlength=n)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(counts, 0)
    return indptr, terms, vals


def columns(n: int, seed: int, device, n_scopes: int = N_SCOPES):
    """scope int32 [n], created / modified int64 [n] on the device."""
    g = torch.Generator(device=device).manual_seed(555 + seed)
    cdf = _zipf_cdf(n_scopes, 1.0, device)
    scope = torch.searchsorted(cdf, torch.rand(n, generator=g, device=device, dtype=torch.float64)).clamp_(max=n_scopes - 1)
    modified = torch.randint(T0, T1, (n,), generator=g, device=device, dtype=torch.int64)
    created = modified - torch.randint(0, 10_000_000, (n,), generator=g, device=device, dtype=torch.int64)
    missing = torch.rand(n, generator=g, device=device) < 0.05
    modified = torch.where(missing, torch.full_like(modified, TS_MISSING), modified)
    return scope.to(torch.int32), created, modified


def queries(nq: int, seed: int, rows_bf16: torch.Tensor, indptr: torch.Tensor | None, terms: torch.Tensor | None,
            nnz=(3, 12), noise: float = 0.5):
    """Host-side query batch: dense float32 [nq, dim] (unit norm) and sparse [(indices, values)]."""
    rng = np.random.RandomState(4321 + seed)
    n, dim = rows_bf16.shape
    pick = rng.randint(0, n, size=nq)
    base = rows_bf16[torch.from_numpy(pick).to(rows_bf16.device)].float().cpu().numpy()
    base /= np.maximum(np.linalg.norm(base, axis=1, keepdims=True), 1e-9)
    q = base + noise * rng.randn(nq, dim).astype(np.float32) / np.sqrt(dim)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    sparse = None
    if indptr is not None:
        sparse = []
        ip = indptr.cpu().numpy() if indptr.numel() <= (1 << 26) else None
        for r in pick:
            lo, hi = (int(ip[r]), int(ip[r + 1])) if ip is not None else (int(indptr[r]), int(indptr[r + 1]))
            t = terms[lo:hi].cpu().numpy().astype(np.int64)
            k = min(len(t), int(rng.randint(nnz[0], nnz[1] + 1)))
            sel = list(rng.choice(t, size=k, replace=False)) if k else []
            extra = int(hash_term(torch.tensor(int(rng.randint(0, VOCAB)))).item())
            if extra not in sel:
                sel.append(extra)
            sparse.append(([int(x) for x in sel], [1.0] * len(sel)))
    return q, sparse


def scope_filter(scope: torch.Tensor, target: float, seed: int, n_scopes: int = N_SCOPES) -> np.ndarray:
    """Random set of scope ids whose row mass is ~target; returns the uint32 bitset."""
    rng = np.random.RandomState(31337 + seed)
    mass = torch.bincount(scope.long(), minlength=n_scopes).cpu().numpy().astype(np.float64)
    mass /= mass.sum()
    order = rng.permutation(n_scopes)
    bits = np.zeros((n_scopes + 31) // 32, np.uint32)
    acc = 0.0
    for s in order:
        if acc + mass[s] > target * 1.02:
            continue
        bits[s >> 5] |= np.uint32(1 << (s & 31))
        acc += mass[s]
        if acc >= target * 0.98:
            break
    return bits
