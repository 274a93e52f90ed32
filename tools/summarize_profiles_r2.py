#!/usr/bin/env python3
"""Turn the artefacts of tools/gpu_final_r2.sh (gpurun_out/r02_*) into the tracked summaries under profiles/:
ncu summary per capture, per-launch time tables, bench lines, sanitizer summaries, traffic.json.
Usage: python tools/summarize_profiles_r2.py [r02]"""
import csv
import json
import shutil
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
# capture name -> (workload, kernel, rows per GPU, queries per batch)
CAPS = {"k2t": ("cfg4", "vb_dense_gemm_tiled_kernel", 12_500_000, 1024), "k3m": ("cfg4", "vb_ms_score_kernel", 12_500_000, 1024),
        "mask": ("cfg4", "vb_mask_kernel", 12_500_000, 1024), "compact": ("cfg4", "vb_compact_kernel", 12_500_000, 1024),
        "k1": ("cfg3-b1-s50", "vb_dense_scan_kernel", 10_000_000, 1), "k1f": ("cfg1", "vb_dense_scan1_kernel", 100_000, 1),
        "k3": ("cfg5-shard", "vb_sparse_kernel", 6_250_000, 4096), "k2": ("cfg2", "vb_dense_gemm_kernel", 1_000_000, 64),
        "sel": ("cfg4", "vb_rowsel_gather_kernel", 12_500_000, 1024)}


def raw(rep):
    c = rep.with_suffix(".csv")
    if c.exists() and c.stat().st_size:
        txt = c.read_text()
    else:
        txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return rows[0], rows[1], rows[2:]


def num(v, unit):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit.lower(), 1)


def main():
    PROF.mkdir(exist_ok=True)
    lines, traffic = [], []
    for name, (wl, kern, rows_gpu, qpb) in CAPS.items():
        rep = OUT / f"{R}_prof_{name}.ncu-rep"
        if not rep.exists() and not rep.with_suffix(".csv").exists():
            continue
        hdr, units, data = raw(rep)
        if not data:
            continue
        idx = {h: i for i, h in enumerate(hdr)}
        d = data[0]
        lines.append(f"== {rep.name}: workload {wl}, ncu --set full --clock-control none (cold cache, serialised) ==")
        lines.append(f"kernel: {d[idx['Kernel Name']][:100]}")
        for k in KEYS:
            if k in idx:
                lines.append(f"  {k:72s} {d[idx[k]]:>18s} {units[idx[k]]}")
        st = [(h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), float(d[idx[h]].replace(",", "") or 0))
              for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        lines.append("  warp stalls per issue: " + ", ".join(f"{n_} {v:.2f}" for n_, v in sorted(st, key=lambda x: -x[1])[:7]))
        lines.append("")
        try:
            traffic.append({"workload": wl, "kernel": kern, "rows_per_gpu": rows_gpu, "queries_per_batch": qpb,
                            "dram_bytes_per_launch": int(num(d[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) +
                                                         num(d[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])),
                            "launch_us": num(d[idx["gpu__time_duration.sum"]], "byte") * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(units[idx["gpu__time_duration.sum"]].lower(), 1),
                            "capture": rep.name})
        except Exception as e:
            lines.append(f"  (traffic not extracted: {e})")
    if lines:
        (PROF / f"{R}_ncu_summary.txt").write_text("\n".join(lines) + "\n")
    if traffic:
        (PROF / "traffic.json").write_text(json.dumps({"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the ncu --set full "
                                                        "captures summarised in %s_ncu_summary.txt; bench.py prints a traffic figure only when workload, "
                                                        "rows per GPU, batch and kernel all match a capture" % R, "captures": traffic}, indent=1) + "\n")
    for lc in sorted(OUT.glob(f"{R}_launches_*.csv")):
        rows = list(csv.DictReader([l for l in lc.read_text().splitlines() if not l.startswith("==")]))
        if not rows:
            continue
        agg = defaultdict(list)
        for r in rows:
            try:
                agg[r["Kernel Name"].split("(")[0]].append(float(r["Metric Value"]) / 1e3)
            except Exception:
                pass
        total = sum(sum(v) for v in agg.values())
        out = [f"# {lc.name}: bench.py --workload {lc.stem.split('launches_')[1]} under ncu --metrics gpu__time_duration.sum --clock-control none, first {len(rows)} search launches",
               "# (per-launch times are cold-cache and serialised: compare SHARES, not absolutes)",
               f"# {'kernel':50s} {'launches':>8s} {'total_us':>10s} {'mean_us':>9s} {'share':>7s}"]
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            out.append(f"  {k[:50]:50s} {len(v):8d} {sum(v):10.1f} {sum(v) / len(v):9.2f} {100 * sum(v) / total:6.1f}%")
        seq = [(r["Kernel Name"].split("(")[0][:48], float(r["Metric Value"]) / 1e3, r.get("Grid Size", "")) for r in rows]
        starts = [i for i, x in enumerate(seq) if x[0].startswith("vb_init_lists")]
        if len(starts) < 2:                                 # (the list set-up is folded into the query prep on most paths)
            starts = [i for i, x in enumerate(seq) if x[0].startswith("vb_prep_query")]
        if len(starts) >= 2:
            out.append("# one batch, launch by launch (us, grid):")
            out += [f"  {n:48s} {t:10.2f} {g}" for n, t, g in seq[starts[-2]:starts[-1]]]
        (PROF / f"{lc.stem}.txt").write_text("\n".join(out) + "\n")
    for f in sorted(OUT.glob(f"{R}_bench_*.json")) + [OUT / f"{R}_pytest_gpu.log", OUT / f"{R}_smoke.log"]:
        if f.exists() and f.stat().st_size:
            shutil.copy(f, PROF / f.name)
    f = OUT / f"{R}_pytest_bounds.log"
    if f.exists():
        (PROF / f"{R}_pytest_bounds.txt").write_text("GPU suite on the bounds-checked build (libvoitta_b200_dbg.so, every VB_CHECK a device-side assert):\n" + "\n".join(f.read_text().splitlines()[-3:]) + "\n")
    f = OUT / f"{R}_sanitize_version.log"
    if f.exists():
        (PROF / f"{R}_sanitize_version.txt").write_text(f.read_text()[:2000])
    for tool in ("plain", "bounds", "memcheck", "racecheck", "synccheck"):
        f = OUT / f"{R}_sanitize_{tool}.log"
        if f.exists():
            keep = [l for l in f.read_text().splitlines() if any(k in l for k in ("SUMMARY", "ok ", "done", "Error", "hazard", "ERROR", "COMPUTE-SANITIZER"))]
            (PROF / f"{R}_sanitize_{tool}.txt").write_text("\n".join(keep[:200]) + "\n")
    print("\n".join(lines[:80]))


if __name__ == "__main__":
    main()
