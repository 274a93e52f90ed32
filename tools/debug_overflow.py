#!/usr/bin/env python3
"""Which candidate list overflows?  Builds a bench workload's shard, runs its query batches through the staged API and,
when vb_fetch reports an overflow, prints the list (dense / sparse), the query's sparse terms and their document
frequencies.  Usage: python tools/debug_overflow.py cfg5-shard"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch                                            # noqa: E402
import bench                                            # noqa: E402
from voitta_rag_b200 import engine, synth               # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5-shard"
cfg = dict(bench.WORKLOADS[name])
device = torch.device("cuda", 0)
ix, keep, (lo, hi), _ = bench.build_shard(cfg, 0, 1, device, torch, synth, engine)
batches, flt = bench.make_batches(cfg, keep, 1, synth, engine, torch)
B, limit = cfg["batch"], cfg["limit"]
filters = None if flt is None else [engine.Filter(*flt)]
fo = None if flt is None else np.zeros(B, np.int32)
n_over = 0
for i, (q, sp) in enumerate(batches):
    st = ix.stage(q, sp, filters, fo, limit=limit, kprime=3 * limit, fusion=cfg["fusion"])
    ix.run_local(None)
    ix.run_fuse(0, None)
    res = ix.fetch(st, allow_overflow=True)
    s = ix.stats()
    if res is None:
        n_over += 1
        li = int(s["last_overflow_first"])
        qi = li % B
        terms = np.asarray(sp[qi][0], np.uint32)
        df, n_live = ix.term_stats(terms)
        print(f"batch {i}: {s['last_overflow_lists']} list(s) overflowed; first = {'sparse' if li >= B else 'dense'} list of query {qi}: "
              f"{len(terms)} terms, df = {sorted(int(x) for x in df)} of {n_live} rows", flush=True)
    else:
        print(f"batch {i}: ok ({s['last_launches']} launches, {s['last_search_ms']:.2f} ms)", flush=True)
print("overflowing batches:", n_over)
