#!/usr/bin/env python3
"""bench.py — hybrid filtered top-k queries/sec on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl b200|reference]

One "step" = one pass of the hot path over one batch of synthetic queries.
  value   whole-job queries/s with the batch already resident in HBM (vb_stage done before the
          timed region): CUDA events on the launching stream around vb_run_local (+ all-gather)
          + vb_run_fuse, max over ranks.
  e2e     the same metric through the C-ABI call with HOST buffers (vb_search: staging, H2D, kernels,
          D2H, decode inside the timed region).
  roofline  dominant kernel: algorithmic bytes per step / its CUDA-event duration, against
          MEASURED_PEAKS.json.
  cpu_baseline  the C oracle (port of the reference's CPU algorithm) on the host cores, one batch.
Multi-GPU (torchrun, one rank per GPU): the corpus is row-sharded over the ranks, the batch grows
with N (per-GPU work fixed => "weak"), candidates are exchanged by one NCCL all-gather per step.
`--impl reference` times the CPU port alone (rank 0), same config/metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # BASELINE.json configs[1] — the configuration the metric is quoted on at 1 GPU
    "cfg2": dict(n=1_000_000, dim=384, batch=64, limit=10, fusion="rrf", sel=None, dist="C",
                 desc="1M chunks x 384-d hybrid dense+sparse (BM25-style) RRF, batch 64, top-10"),
    # BASELINE.json configs[0] shape (the reference's CPU-runnable case), dense only
    "cfg1": dict(n=100_000, dim=384, batch=1, limit=10, fusion="dense", sel=None, dist="C",
                 desc="100k chunks x 384-d dense cosine top-10, single query"),
    # BASELINE.json configs[2]
    "cfg3-b1-s50": dict(n=10_000_000, dim=768, batch=1, limit=10, fusion="rrf", sel=0.5, dist="C",
                        desc="10M x 768-d hybrid, scope+time filter 50%, batch 1"),
    "cfg3-b1-s1": dict(n=10_000_000, dim=768, batch=1, limit=10, fusion="rrf", sel=0.01, dist="C",
                       desc="10M x 768-d hybrid, scope+time filter 1%, batch 1"),
    "cfg3-b256-s50": dict(n=10_000_000, dim=768, batch=256, limit=10, fusion="rrf", sel=0.5, dist="C",
                          desc="10M x 768-d hybrid, scope+time filter 50%, batch 256"),
    # BASELINE.json configs[3] as ONE of its 8 row shards (100M / 8 rows per GPU, the full batch): the per-GPU
    # work of the 8-GPU run; needs --no-cpu-baseline (the fp32 host copy would not fit next to it)
    "cfg4-shard": dict(n=12_500_000, dim=768, batch=1024, limit=100, fusion="rrf", sel=0.5, dist="C",
                       desc="one 12.5M-row shard of 100M x 768-d hybrid, filter 50%, batch 1024, top-100"),
    # BASELINE.json configs[3] itself when run with --gpus 8 (12.5M rows per GPU = 100M rows, batch 1024 for the whole
    # job, top-100); at --gpus 2/4 the same per-GPU shard size with 25M/50M rows in total
    "cfg4": dict(rows_per_gpu=12_500_000, n=12_500_000, dim=768, batch=1024, fixed_batch=True, limit=100, fusion="rrf", sel=0.5, dist="C",
                 desc="100M x 768-d hybrid at 8 GPUs (12.5M rows per GPU), filter 50%, batch 1024, top-100"),
    # BASELINE.json configs[4] as one of its 8 row shards: MCP replay, 4096 mixed-length queries (2..64 sparse terms),
    # 50M x 1024-d / 8 = 6.25M rows per GPU, filtered top-20
    "cfg5-shard": dict(n=6_250_000, dim=1024, batch=4096, limit=20, fusion="rrf", sel=0.5, dist="C", qnnz=(2, 64),
                       desc="one 6.25M-row shard of 50M x 1024-d MCP replay, 4096 mixed-length queries, filter 50%, top-20"),
    # the per-GPU work of cfg2 at 8 GPUs (125k rows, batch 512): for profiling the weak-scaling overheads on one GPU
    "cfg2-g8shard": dict(n=125_000, dim=384, batch=512, limit=10, fusion="rrf", sel=None, dist="C",
                         desc="one 125k-row shard of cfg2 at 8 GPUs, batch 512, top-10"),
    # small shape for quick checks
    "tiny": dict(n=65_536, dim=128, batch=16, limit=10, fusion="rrf", sel=None, dist="C", desc="tiny smoke shape"),
}
BLOCK_ROWS = 125_000        # corpus generation granule: shard boundaries at 1/2/4/8 GPUs coincide with it
N_QUERY_BATCHES = 8


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        inside = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], None, set()
        for r in inside:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_shard(cfg, rank, world, device, torch, synth, engine):
    """Generate this rank's rows on the device and load them into an Index."""
    n = cfg["n"]
    per = (n + world - 1) // world
    lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
    ix = engine.Index(cfg["dim"], device=device.index or 0, row_base=lo)
    keep = None
    r = lo
    while r < hi:
        blk = r // BLOCK_ROWS
        b_lo, b_hi = blk * BLOCK_ROWS, min(n, (blk + 1) * BLOCK_ROWS)
        rows = synth.dense_rows(b_hi - b_lo, cfg["dim"], blk, device, cfg["dist"])
        ip, tm, vl = synth.sparse_rows(b_hi - b_lo, blk, device)
        sc, cr, mo = synth.columns(b_hi - b_lo, blk, device)
        s, e = r - b_lo, min(hi, b_hi) - b_lo
        ipc = (ip[s:e + 1] - ip[s]).contiguous()
        a, z = int(ip[s]), int(ip[e])
        rs, tms, vls = rows[s:e].contiguous(), tm[a:z].contiguous(), vl[a:z].contiguous()
        scs, crs, mos = sc[s:e].contiguous(), cr[s:e].contiguous(), mo[s:e].contiguous()
        torch.cuda.synchronize(device)
        ix.upsert_dev(e - s, rs.data_ptr(), ipc.data_ptr(), tms.data_ptr(), vls.data_ptr(), scs.data_ptr(),
                      crs.data_ptr(), mos.data_ptr())
        if keep is None:
            keep = dict(rows=rs, ip=ipc, tm=tms, scope=scs, lo=r)
        r = b_lo + e
    return ix, keep, (lo, hi)


def make_batches(cfg, keep, world, synth, engine, torch):
    """Rank-0 query batches (dense fp32 + sparse terms) and, if the workload filters, one filter."""
    B = cfg["batch"] * (1 if cfg.get("fixed_batch") else world)
    batches = []
    for i in range(N_QUERY_BATCHES):
        q, sp = synth.queries(B, i, keep["rows"], keep["ip"], keep["tm"], nnz=cfg.get("qnnz", (3, 12)))
        batches.append((q, sp if cfg["fusion"] != "dense" else None))
    flt = None
    if cfg["sel"] is not None:
        # folder scope mask with half of the selectivity budget, time range with the rest
        s = float(cfg["sel"]) ** 0.5
        bits = synth.scope_filter(keep["scope"], s, 0)
        span = synth.T1 - synth.T0
        lo = synth.T0 + int(0.1 * span)
        hi = lo + int(span * s / 0.95)
        flt = (bits, engine.TS_MODIFIED, lo, hi)
    return batches, flt


def run_reference(args, cfg):
    """--impl reference: the CPU port of the reference's algorithm on the host cores, alone."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from voitta_rag_b200 import synth
    from oracle import oracle_c
    device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    cc, batches, flt, threads = host_corpus(cfg, device, torch, synth, oracle_c, world=args.gpus)
    sample = min(cfg["batch"] * args.gpus, 16)
    fz = {"dense": 0, "weighted": 1, "rrf": 2}[cfg["fusion"]]

    def step(i):
        q, sp = batches[i % len(batches)]
        cc.search_batch(q[:sample], None if sp is None else sp[:sample], None if flt is None else [flt],
                        None if flt is None else np.zeros(sample, np.int32), limit=cfg["limit"], fusion=fz)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "hybrid filtered top-k queries/sec", "value": v, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {cfg['desc']}", "queries_per_step": sample},
        "cpu_baseline": {"value": v, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} queries of the batch per step over the full {cfg['n']}-row corpus"},
        "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def host_corpus(cfg, device, torch, synth, oracle_c, world=1):
    """The full corpus as host arrays for the C oracle (fp32 copies of the bf16 rows)."""
    n, dim = cfg["n"], cfg["dim"]
    dense = np.empty((n, dim), np.float32)
    ips, tms, vls, scs, crs, mos = [np.zeros(1, np.int64)], [], [], [], [], []
    keep = None
    for blk in range((n + BLOCK_ROWS - 1) // BLOCK_ROWS):
        lo, hi = blk * BLOCK_ROWS, min(n, (blk + 1) * BLOCK_ROWS)
        rows = synth.dense_rows(hi - lo, dim, blk, device, cfg["dist"])
        ip, tm, vl = synth.sparse_rows(hi - lo, blk, device)
        sc, cr, mo = synth.columns(hi - lo, blk, device)
        dense[lo:hi] = rows.float().cpu().numpy()
        ips.append(ip[1:].cpu().numpy() + ips[-1][-1])
        tms.append(tm.cpu().numpy().astype(np.uint32)); vls.append(vl.cpu().numpy())
        scs.append(sc.cpu().numpy().astype(np.uint32)); crs.append(cr.cpu().numpy()); mos.append(mo.cpu().numpy())
        if keep is None:
            keep = dict(rows=rows, ip=ip, tm=tm, scope=sc, lo=0)
    from voitta_rag_b200 import engine
    batches, flt = make_batches(cfg, keep, world, synth, engine, torch)
    cc = oracle_c.CorpusC(dense, (np.concatenate(ips), np.concatenate(tms), np.concatenate(vls)),
                          np.concatenate(scs), np.concatenate(crs), np.concatenate(mos))
    return cc, batches, flt, oracle_c.num_threads()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dense-path", type=int, default=0, help="0 auto, 1 GEMV scan, 2 tcgen05 GEMM")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = dict(WORKLOADS[args.workload])
    if "rows_per_gpu" in cfg:
        cfg["n"] = cfg["rows_per_gpu"] * max(1, args.gpus)
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args, cfg)

    import torch
    import torch.distributed as dist
    from voitta_rag_b200 import engine, synth
    from voitta_rag_b200.sharded import ShardedIndex

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 backend has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus}")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    ix, keep, (lo, hi) = build_shard(cfg, rank, world, device, torch, synth, engine)
    if args.dense_path:
        ix.set_option("dense_path", args.dense_path)
    sh = ShardedIndex(ix, rank, world, device=device)
    obj = [None]
    if rank == 0:
        obj[0] = make_batches(cfg, keep, world, synth, engine, torch)
    if world > 1:
        dist.broadcast_object_list(obj, src=0)
    batches, flt = obj[0]
    B = cfg["batch"] * (1 if cfg.get("fixed_batch") else world)
    limit = cfg["limit"]
    hybrid = cfg["fusion"] != "dense"
    kprime = limit * 3 if hybrid else limit
    filters = None if flt is None else [engine.Filter(*flt)]
    filter_of = None if flt is None else np.zeros(B, np.int32)
    if hybrid:
        sh.finalize_from_queries([sp for _, sp in batches])
    weighted = [sh.idf_weights(sp) if hybrid else None for _, sp in batches]

    def stage(i):
        q, _ = batches[i % len(batches)]
        return ix.stage(q, weighted[i % len(batches)], filters, filter_of, limit=limit, kprime=kprime,
                        fusion=cfg["fusion"], sparse_weight=0.1, apply_idf=False)

    def device_step(staged):
        """The timed device work of one step; returns the result (fetch = D2H, outside the events)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(sh.stream):
            ev0.record(sh.stream)
            if world == 1:
                ix.run_local(None)
                ix.run_fuse(0, None)
            else:
                sh._enqueue(B, kprime)      # threshold all-reduce, local branches, candidate all-gather, merge + fuse
            ev1.record(sh.stream)
        res = ix.fetch(staged, allow_overflow=True)
        if res is None:
            raise SystemExit("bench.py: candidate overflow in the timed region (unexpected for this workload)")
        return ev0.elapsed_time(ev1), res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- device-resident timing ("value") ----------------------------------------------------------
    for i in range(args.warmup):
        device_step(stage(i))
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t_begin = time.perf_counter()
    ms = np.zeros(args.steps)
    launches = 0
    for i in range(args.steps):
        st = stage(args.warmup + i)                 # host staging + H2D: outside the timed events
        if world > 1:                               # ranks enter the step together: otherwise the all-gather of the
            torch.cuda.synchronize(device)          # faster rank waits for the other's host staging inside the events
            dist.barrier()
        ms[i], _ = device_step(st)
        launches += ix.stats()["last_launches"]
    barrier()
    t_end = time.perf_counter()
    if world > 1:
        t = torch.from_numpy(ms).to(device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.cpu().numpy()
    clocks = sampler.stop(t_begin, t_end)
    total_ms = float(ms.sum())
    value = B * args.steps / (total_ms / 1e3)
    dense_path = ix.stats()["last_dense_path"]

    # ---- end to end through the C ABI with host buffers -------------------------------------------
    # host buffers (numpy arrays + C structs) are the call's inputs; they are built once per batch
    if world == 1:
        packed = [ix.pack(q, sp, filters, filter_of, limit=limit, kprime=kprime, fusion=cfg["fusion"], sparse_weight=0.1)
                  for q, sp in batches]
    else:
        packed = [sh.pack(q, sp, filters, filter_of, limit=limit, kprime=kprime, fusion=cfg["fusion"], sparse_weight=0.1)
                  for q, sp in batches]

    def e2e_step(i):
        if world == 1:
            return ix.search_packed(packed[i % len(batches)])
        return sh.search_packed(packed[i % len(batches)])

    for i in range(args.warmup):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    if world == 1:
        # a stream of batches, two in flight: batch i+1 is staged (host prep + H2D) while batch i runs;
        # every batch has its own H2D copy from pinned memory and its own D2H read of the results
        n_done = 0
        for last in ix.search_stream(packed[(args.warmup + i) % len(batches)] for i in range(args.steps)):
            n_done += 1
        assert n_done == args.steps
    else:
        n_done = 0
        for last in sh.search_stream(packed[(args.warmup + i) % len(batches)] for i in range(args.steps)):
            n_done += 1
        assert n_done == args.steps
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    st = ix.stats()
    e2e = {"value": B * args.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(st["last_h2d_bytes"]),
           "d2h_bytes_per_step": int(st["last_d2h_bytes"]), "ms_per_step": 1e3 * e2e_s / args.steps}

    # ---- per-kernel durations: a short pass with the two chains serialised on one stream and CUDA
    # events around every phase (the timed regions above run with overlap on and no phase events) ----
    ix.set_option("profile", 1)
    ix.set_option("overlap", 0)
    phase = np.zeros(5)
    big = np.zeros(2)
    n_prof = max(3, min(args.steps, 50))
    for i in range(n_prof):
        device_step(stage(i))
        s_ = ix.stats()
        phase += [s_["last_mask_ms"], s_["last_dense_ms"], s_["last_sparse_ms"], s_["last_select_ms"], s_["last_fuse_ms"]]
        big += [s_["last_dense_big_ms"], s_["last_sparse_big_ms"]]
    big_rows = int(ix.stats()["last_big_rows"])
    ix.set_option("profile", 0)
    ix.set_option("overlap", 1)

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    hbm_peak, tf_peak, peak_kind = peaks()
    rows_local = hi - lo
    names = ["mask", "dense", "sparse", "select", "fuse"]
    per_step = phase / n_prof
    dom = int(np.argmax(per_step))
    d_pad = (cfg["dim"] + 63) // 64 * 64
    passes = max(1, int(ix.stats()["last_dense_passes"]))     # corpus passes of the dense kernel per step
    sel = cfg["sel"] if cfg["sel"] is not None else 1.0
    dense_bytes = passes * rows_local * ((d_pad * 2 + 4) * (sel if dense_path == 1 else 1.0)) + (rows_local / 8 if flt else 0)
    dense_flops = 2.0 * B * rows_local * d_pad
    sparse_bytes = 0.0
    if hybrid:
        for _, sp in batches:
            terms = np.asarray([t for s_ in sp for t in s_[0]], np.uint32)
            df, _ = ix.term_stats(terms)
            sparse_bytes += float(df.sum()) * 8
        sparse_bytes /= len(batches)
    alg = {"mask": rows_local * (4 + 8) + rows_local / 8, "dense": dense_bytes, "sparse": sparse_bytes,
           "select": 0.0, "fuse": 0.0}
    dom_name = names[dom]
    # the roofline is quoted on ONE launch: the dominant kernel's launch over the largest segment
    # (the last one, `big_rows` of the shard's rows), whose ncu capture is in profiles/
    frac_rows = big_rows / max(1, rows_local)
    if dom_name in ("dense", "sparse") and big[0 if dom_name == "dense" else 1] > 0:
        launch_ms = float(big[0 if dom_name == "dense" else 1] / n_prof)
        launch_bytes = (alg["dense"] / passes if dom_name == "dense" else alg["sparse"]) * frac_rows
    else:
        launch_ms, launch_bytes = float(per_step[dom]), alg[dom_name]
    achieved = launch_bytes / (launch_ms / 1e3) / 1e9 if launch_ms > 0 else 0.0
    kname = {"dense": "vb_dense_gemm_kernel" if dense_path == 2 else "vb_dense_scan_kernel", "sparse": "vb_sparse_kernel",
             "mask": "vb_mask_kernel", "select": "vb_compact_kernel", "fuse": "vb_fuse_kernel"}
    traffic = None
    tfile = ROOT / "profiles" / "traffic.json"          # dram bytes per launch from the committed ncu --set full capture
    if tfile.exists():
        t = json.loads(tfile.read_text()).get(args.workload, {}).get(kname[dom_name])
        traffic = t["dram_bytes_per_launch"] if t else None
    tf_launch = (dense_flops * frac_rows / (big[0] / n_prof / 1e3) / 1e12) if big[0] > 0 else None
    tensor_bound = dom_name == "dense" and dense_path == 2 and B / passes > 250      # past the ridge (252 flop/B)
    roofline = {"kernel": kname[dom_name] if not tensor_bound else "vb_dense_gemm_tiled_kernel",
                "launch": "largest segment: %d of %d rows%s" % (
                    big_rows, rows_local, "" if dom_name != "dense" else ", one of %d pass(es)" % passes),
                "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind}, burst copy)", "traffic": traffic,
                "algorithmic_bytes_per_launch": launch_bytes, "kernel_ms_per_launch": launch_ms,
                "how": "CUDA events around each kernel on its launching stream, chains serialised on one stream, %d steps after the timed region" % n_prof,
                "dense_tflops_per_step": dense_flops / (per_step[1] / 1e3) / 1e12 if per_step[1] > 0 else None,
                "phase_ms_per_step": {n_: float(v) for n_, v in zip(names, per_step)},
                "phase_gbs_per_step": {n_: (alg[n_] / (per_step[j] / 1e3) / 1e9 if per_step[j] > 0 else None) for j, n_ in enumerate(names)},
                "big_launch": {"dense_ms": float(big[0] / n_prof), "sparse_ms": float(big[1] / n_prof),
                               "dense_gbs": (alg["dense"] / passes * frac_rows) / (big[0] / n_prof / 1e3) / 1e9 if big[0] > 0 else None,
                               "dense_tflops": tf_launch,
                               "sparse_gbs": (alg["sparse"] * frac_rows) / (big[1] / n_prof / 1e3) / 1e9 if big[1] > 0 else None}}
    if tensor_bound and tf_launch:
        # batched scoring past the ridge point: the tensor pipe is the roofline (flops = 2 * rows * B * d_pad)
        roofline.update({"bound": "tensor", "achieved": tf_launch, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf_launch / tf_peak,
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops ({peak_kind}, cuBLAS burst)",
                         "algorithmic_flops_per_launch": dense_flops * frac_rows, "hbm_gbs_same_launch": achieved})

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on one batch, plus a parity spot check ----
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle_c
        cc, _, _, threads = host_corpus(cfg, device, torch, synth, oracle_c)
        q, sp = batches[0]
        nq = min(B, 64)
        fz = {"dense": 0, "weighted": 1, "rrf": 2}[cfg["fusion"]]
        t0 = time.perf_counter()
        want = cc.search_batch(q[:nq], None if sp is None else sp[:nq], None if flt is None else [flt],
                               None if flt is None else np.zeros(nq, np.int32), limit=limit, kprime=kprime, fusion=fz)
        dt = time.perf_counter() - t0
        cpu = {"value": nq / dt, "unit": "queries/s", "cores": threads, "kind": "port",
               "sample": f"{nq} queries (one batch) over the full {cfg['n']}-row corpus, oracle/oracle_c.c with OpenMP"}
        got = ix.search_batch(q[:nq], None if sp is None else sp[:nq], filters, None if flt is None else np.zeros(nq, np.int32),
                              limit=limit, kprime=kprime, fusion=cfg["fusion"], branches=True)
        same = sum(int(np.array_equal(got.rows[i, :got.counts[i]], want["rows"][i, :want["counts"][i]])) for i in range(nq))
        same_d = sum(int(np.array_equal(got.dense_rows[i, :got.dense_counts[i]], want["dense_rows"][i, :want["dense_counts"][i]])) for i in range(nq))
        same_s = sum(int(np.array_equal(got.sparse_rows[i, :got.sparse_counts[i]], want["sparse_rows"][i, :want["sparse_counts"][i]])) for i in range(nq))
        parity = {"queries": nq, "fused_identical": same, "dense_branch_identical": same_d, "sparse_branch_identical": same_s,
                  "note": "bf16 query rounding on the tensor-core path may swap near-ties (<=1e-3 rel)"}

    if rank == 0:
        line = {
            "metric": "hybrid filtered top-k queries/sec", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {cfg['desc']}", "rows_total": cfg["n"], "rows_per_gpu": rows_local,
                       "dim": cfg["dim"], "queries_per_step": B, "limit": limit, "kprime": kprime, "fusion": cfg["fusion"],
                       "selectivity": cfg["sel"], "dense_path": {1: "K1 GEMV scan", 2: "K2 tcgen05 GEMM"}.get(dense_path),
                       "parallelism": (f"row-sharded x{world}, batch {B} for the whole job, NCCL all-gather of candidates" if cfg.get("fixed_batch") else f"row-sharded x{world}, batch {cfg['batch']}/GPU, NCCL all-gather of candidates") if world > 1 else "1 GPU",
                       "l2": "corpus per GPU (%.0f MB bf16 + postings) exceeds the 126 MB L2; %d distinct query batches rotate" % (rows_local * d_pad * 2 / 1e6, len(batches))},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "parity_spot_check": parity,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    ix.close()


if __name__ == "__main__":
    main()
