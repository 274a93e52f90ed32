#!/usr/bin/env python3
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/<round>_*) into the small, tracked
summaries under profiles/.  Usage: python tools/summarize_profiles.py r01"""
import csv
import json
import shutil
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    PROF.mkdir(exist_ok=True)
    lines = []
    for name in ("gemm", "sparse", "tiled"):
        rep = OUT / f"{R}_prof_{name}.ncu-rep"
        if not rep.exists():
            continue
        hdr, units, data = raw(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        lines.append(f"== {rep.name}: ncu --set full --clock-control none (cold cache, serialised) ==")
        for d in data:
            lines.append(f"kernel: {d[idx['Kernel Name']][:90]}")
            for k in KEYS:
                if k in idx:
                    lines.append(f"  {k:72s} {d[idx[k]]:>16s} {units[idx[k]]}")
            stalls = [(h, float(d[idx[h]].replace(',', '') or 0)) for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
            tot = sum(v for _, v in stalls) or 1
            lines.append("  warp stall samples: " + ", ".join(f"{h.split('stalled_')[1]} {100 * v / tot:.0f}%" for h, v in sorted(stalls, key=lambda x: -x[1])[:7]))
        lines.append("")
    (PROF / f"{R}_ncu_summary.txt").write_text("\n".join(lines))
    lc = OUT / f"{R}_launches_cfg2.csv"
    if lc.exists():
        rows = list(csv.DictReader([l for l in lc.read_text().splitlines() if not l.startswith("==")]))
        agg = defaultdict(list)
        for r in rows:
            agg[r["Kernel Name"].split("(")[0]].append(float(r["Metric Value"]) / 1e3)
        total = sum(sum(v) for k, v in agg.items() if not any(x in k for x in ("ingest", "posting", "set_alive", "offset_i64", "find_tail")))
        out = [f"# {lc.name}: python bench.py --steps 3 --warmup 3 --no-cpu-baseline under",
               "# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vb_  (per-launch times are cold-cache and serialised: compare SHARES)",
               f"# {'kernel':50s} {'launches':>8s} {'total_us':>10s} {'mean_us':>9s} {'share_of_search':>16s}"]
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            search = not any(x in k for x in ("ingest", "posting", "set_alive", "offset_i64", "find_tail"))
            out.append(f"  {k[:50]:50s} {len(v):8d} {sum(v):10.1f} {sum(v) / len(v):9.2f} {(100 * sum(v) / total if search else 0):15.1f}%")
        seq = [(r["Kernel Name"].split("(")[0][:44], float(r["Metric Value"]) / 1e3, r["Grid Size"]) for r in rows]
        starts = [i for i, x in enumerate(seq) if x[0].startswith("vb_init_lists")]
        if len(starts) >= 3:
            out.append("# one search step, launch by launch (us, grid):")
            out += [f"  {n:44s} {t:9.2f} {g}" for n, t, g in seq[starts[-2]:starts[-1]]]
        (PROF / f"{R}_launches_cfg2.txt").write_text("\n".join(out) + "\n")
    for f in sorted(OUT.glob(f"{R}_bench_*.json")) + [OUT / f"{R}_pytest_gpu.log", OUT / f"{R}_smoke.log"]:
        if f.exists() and f.stat().st_size:
            shutil.copy(f, PROF / f.name)
    # dram bytes per launch of the captured kernels -> traffic.json (bench.py's roofline.traffic)
    traffic = {}
    for name, wl, kern in (("gemm", "cfg2", "vb_dense_gemm_kernel"), ("sparse", "cfg2", "vb_sparse_kernel"),
                           ("tiled", "cfg3-b256-s50", "vb_dense_gemm_tiled_kernel")):
        rep = OUT / f"{R}_prof_{name}.ncu-rep"
        if not rep.exists():
            continue
        hdr, units, data = raw(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        def val(k):
            v = float(data[0][idx[k]].replace(",", ""))
            u = units[idx[k]].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        traffic.setdefault(wl, {})[kern] = {"dram_bytes_per_launch": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")),
                                            "capture": rep.name}
    if traffic:
        traffic["_comment"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch, from the ncu --set full captures summarised in %s_ncu_summary.txt (largest segment launch of a step)" % R
        (PROF / "traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
