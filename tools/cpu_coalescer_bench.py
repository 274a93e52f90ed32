"""CPU-only check of the host layer's thread behaviour: VectorStoreService.search from 1 and 16 threads over the
test double (tests/_fake_index.py: C oracle under ctypes, GIL released like the real library).  Not a product
benchmark — it shows whether concurrent callers are slower than a lone one (GIL convoy in the coalescer)."""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["EMBEDDING_DIMENSION"] = "64"
os.environ["QDRANT_COLLECTION"] = "coalesce_bench"
from voitta_rag_b200 import vector_store as VS
from _fake_index import FakeIndex

rng = np.random.default_rng(0)
n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 20000, 64
svc = VS.VectorStoreService(_index_factory=lambda: FakeIndex(d))
meta = [VS.ChunkMetadata(file_path=f"root/f{i//40}.md", folder_path=f"root/d{i % 50}", index_folder="root", file_name="f", chunk_index=i % 40, total_chunks=40, start_char=0, end_char=10, indexed_at="2025-01-01T00:00:00") for i in range(n)]
emb = rng.standard_normal((n, d)).astype(np.float32)
sp = [(list(map(int, rng.choice(500, 6, replace=False))), list(map(float, rng.random(6)))) for _ in range(n)]
for i in range(0, n, 5000):
    svc.store_chunks([(f"c{j}", emb[j].tolist(), meta[j]) for j in range(i, min(n, i + 5000))], sparse_vectors=sp[i:i + 5000])
calls = [(rng.standard_normal(d).astype(np.float32).tolist(), sp[i]) for i in range(64)]
inc = [f"root/d{i}" for i in range(25)]
def worker(k, n_calls, lat):
    mine = []
    for j in range(n_calls):
        qe, sq = calls[(k * n_calls + j) % 64]
        t0 = time.perf_counter(); svc.search(qe, limit=20, sparse_query=sq, sparse_weight=0.1, include_folders=inc, scope_key=("u", "p", 1)); mine.append(time.perf_counter() - t0)
    lat.extend(mine)
worker(0, 4, [])
for nt in (1, 4, 16):
    lat = []
    th = [threading.Thread(target=worker, args=(k, 200, lat)) for k in range(nt)]
    t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]; dt = time.perf_counter() - t0
    a = np.sort(np.asarray(lat))
    print(f"threads {nt:2d}: {len(a)/dt:8.0f} q/s  p50 {1e3*a[len(a)//2]:.2f} ms  p99 {1e3*a[int(.99*len(a))]:.2f} ms  {svc.coalescing_stats()}")
