#!/usr/bin/env python3
"""CPU prototype that sized the posting-driven MaxScore kernel (K3M, csrc/sparse_ms.cuh) before any CUDA was
written: on the benchmark's synthetic corpus (voitta_rag_b200.synth) it measures, per query, the posting mass, the
ESSENTIAL postings under the threshold the first rows establish, and how many lookups the progressive bound test
needs.  Results that shaped the design (n rows, k', rows seen before the threshold):
    1M, 30, 65536:   1.09 M postings/query, 26 k essential (2.4 %), 14 k essential lookups, 19 k NE lookups, 402 survivors
    3M, 300, 1M:     2.34 M postings/query, 52 k essential (2.2 %), 28 k + 35 k lookups, 593 survivors
    2M, 60, 500k, mixed-length queries (2..64 terms): 5.75 M postings/query, 254 k essential (4.4 %), 720 k + 266 k lookups
  python tools/proto_maxscore.py <rows> <kprime> <rows_seen> [mixed]
(The prototype orders essential terms by ascending posting count; the kernel ended up with descending ub.)"""
import sys, time, math
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from voitta_rag_b200 import synth
import scipy.sparse as sp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
kp = int(sys.argv[2]) if len(sys.argv) > 2 else 30
seg0 = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
qn = (3, 12) if len(sys.argv) <= 4 else (2, 64)
dev = torch.device('cpu')
t0 = time.time()
ips, tms, vls = [], [], []
base = 0
BLK = 125000
for blk in range((n + BLK - 1) // BLK):
    ip, tm, vl = synth.sparse_rows(min(BLK, n - blk * BLK), blk, dev)
    ips.append(ip[1:].numpy() + base); base = ips[-1][-1]
    tms.append(tm.numpy().astype(np.int64)); vls.append(vl.numpy())
    if blk == 0:
        keep_ip, keep_tm = ip, tm
indptr = np.concatenate([[0]] + ips); terms = np.concatenate(tms); vals = np.concatenate(vls)
print('gen', time.time() - t0, 'nnz', len(terms))
# remap terms to dense ids
uniq, inv = np.unique(terms, return_inverse=True)
T = len(uniq)
M = sp.csr_matrix((vals.astype(np.float64), inv, indptr), shape=(n, T))
Mc = M.tocsc()
df = np.diff(Mc.indptr)
maxval = np.zeros(T); 
maxval = np.maximum.reduceat(Mc.data, Mc.indptr[:-1]) if T else maxval
print('terms', T, 'heavy(df>=n/8):', int((df * 8 >= n).sum()))
rows0 = torch.zeros((BLK, 8), dtype=torch.bfloat16)
NQ = 64
q, sparse = synth.queries(NQ, 0, rows0, keep_ip, keep_tm, nnz=qn)
tot_post = tot_ess = tot_first = tot_surv = 0
tot_heavy_ess = 0
tot_lookups = 0
nclassB = 0
for (tq, vq) in sparse:
    tid = np.searchsorted(uniq, np.asarray(tq)); ok = (tid < T) & (uniq[np.minimum(tid, T - 1)] == np.asarray(tq))
    tid = tid[ok]
    d = df[tid].astype(np.float64)
    w = np.log((n - d + 0.5) / (d + 0.5) + 1.0)
    ub = w * maxval[tid]
    # exact scores over all rows (sparse)
    sc = np.zeros(n)
    touched = np.zeros(n, bool)
    for t, ww in zip(tid, w):
        r = Mc.indices[Mc.indptr[t]:Mc.indptr[t + 1]]; v = Mc.data[Mc.indptr[t]:Mc.indptr[t + 1]]
        sc[r] += ww * v; touched[r] = True
    s0 = np.where(touched[:seg0], sc[:seg0], -np.inf)
    srt = np.sort(s0)[::-1]
    tau = srt[kp - 1] if np.isfinite(srt[kp - 1]) else -np.inf
    # maxscore partition
    order = np.argsort(ub)
    cum = np.cumsum(ub[order])
    ne = order[cum < tau * (1 - 1e-9)] if tau > 0 else order[:0]
    ess = np.setdiff1d(order, ne)
    ubne = ub[ne].sum()
    # postings in segment [seg0, n)
    pall = 0; pess = 0; first = 0
    ub_ess_tot = ub[ess].sum()
    ess_heavy = 0
    cand_rows = {}
    for j in range(len(tid)):
        t = tid[j]
        r = Mc.indices[Mc.indptr[t]:Mc.indptr[t + 1]]; v = Mc.data[Mc.indptr[t]:Mc.indptr[t + 1]]
        m = r >= seg0
        pall += int(m.sum())
        if j in ess:
            pess += int(m.sum())
            if df[t] * 8 >= n: ess_heavy += int(m.sum())
            # first test: w*v + other ess ub + ubne >= tau
            c = w[j] * v[m]
            f = c + (ub_ess_tot - ub[j]) + ubne >= tau
            first += int(f.sum())
    surv = int(((sc[seg0:] > tau) & touched[seg0:]).sum())
    tot_post += pall; tot_ess += pess; tot_first += first; tot_surv += surv; tot_heavy_ess += ess_heavy
    if ess_heavy: nclassB += 1
print(f'n={n} kp={kp} queries={NQ}: postings/q={tot_post/NQ:.0f} essential/q={tot_ess/NQ:.0f} ({100*tot_ess/tot_post:.1f}%) ess_heavy/q={tot_heavy_ess/NQ:.0f} first-pass/q={tot_first/NQ:.0f} survivors/q={tot_surv/NQ:.0f} classB={nclassB}')

# ---- detailed simulation of the posting-driven MaxScore with progressive lookups ----
print('--- progressive lookups ---')
tot_ess = tot_owner = tot_lk_ess = tot_lk_ne = tot_final = 0
for (tq, vq) in sparse:
    tid = np.searchsorted(uniq, np.asarray(tq)); ok = (tid < T) & (uniq[np.minimum(tid, T - 1)] == np.asarray(tq))
    tid = tid[ok]
    d = df[tid].astype(np.float64)
    w = np.log((n - d + 0.5) / (d + 0.5) + 1.0)
    ub = w * maxval[tid]
    sub = Mc[:, tid].tocsr()          # n x nt
    sc = np.asarray(sub @ w).ravel()
    touched = np.diff(sub.indptr) > 0
    s0 = np.where(touched[:seg0], sc[:seg0], -np.inf)
    srt = np.sort(s0)[::-1]
    tau = srt[kp - 1] if np.isfinite(srt[kp - 1]) else -np.inf
    order = np.argsort(ub)
    cum = np.cumsum(ub[order])
    n_ne = int((cum < tau * (1 - 1e-9)).sum()) if tau > 0 else 0
    ne = order[:n_ne]; ess = order[n_ne:]
    # essential sorted by df ascending (rarest first) for ownership
    ess = ess[np.argsort(d[ess])]
    D = sub[seg0:].toarray() * w          # rows x nt contributions (dense small)
    P = D > 0
    # candidate rows: any essential present
    cand = P[:, ess].any(axis=1)
    Dc, Pc = D[cand], P[cand]
    nc = Dc.shape[0]
    tot_ess += int(P[:, ess].sum()); tot_owner += nc
    # owner: lookups of later essential terms = (#ess - 1 - idx of first present)... count each remaining ess term as one lookup
    firstidx = np.argmax(Pc[:, ess], axis=1)
    tot_lk_ess += int((len(ess) - 1 - firstidx).sum())   # lookups into later essential lists
    # (ownership checks for non-owner postings: each posting of term at position p checks p earlier lists)
    partial = Dc[:, ess].sum(axis=1)
    # progressive NE lookups, descending ub
    ne_desc = ne[::-1]
    rem = ub[ne_desc].sum()
    alive = partial + rem >= tau
    lk = 0
    cur = partial.copy()
    for j in ne_desc:
        lk += int(alive.sum())
        cur = cur + Dc[:, j]
        rem -= ub[j]
        alive &= (cur + rem >= tau)
    tot_lk_ne += lk; tot_final += int(alive.sum())
print(f'ess postings/q={tot_ess/NQ:.0f} owners/q={tot_owner/NQ:.0f} ess-lookups/q={tot_lk_ess/NQ:.0f} NE-lookups/q={tot_lk_ne/NQ:.0f} final survivors/q={tot_final/NQ:.0f}')
