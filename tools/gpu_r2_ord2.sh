#!/bin/bash
# Round 2, pass 30: who blocks whom — the copy kernel's and K3M's CTAs per SM (both fill every thread slot by default)
mkdir -p gpurun_out
timeout 900 python tools/ab_opts.py --workload cfg4 --batches 6 --out gpurun_out/r02ord2_ab_cfg4.jsonl --base "overlap=1,sel_ctas=8,ms_ctas=0,dense_wait_sparse=0" \
  --set "sel_ctas=8" --set "sel_ctas=4" --set "sel_ctas=2" --set "sel_ctas=1" --set "sel_ctas=2,ms_ctas=12" --set "sel_ctas=2,ms_ctas=8" --set "sel_ctas=2,ms_ctas=12,dense_wait_sparse=1" \
  --set "sel_ctas=2,ms_ctas=8,dense_wait_sparse=1" --set "sel_ctas=2,dense_wait_sparse=1" --set "sel_ctas=4,ms_ctas=8,dense_wait_sparse=1" 2> gpurun_out/r02ord2_ab_cfg4.err | cut -c1-120,330-700; echo "rc=$?"; tail -2 gpurun_out/r02ord2_ab_cfg4.err
timeout 600 python tools/timeline_dump.py --workload cfg4 --set "sel_ctas=2,ms_ctas=8,dense_wait_sparse=1" --set "sel_ctas=2,ms_ctas=8,dense_wait_sparse=0" > gpurun_out/r02_timeline2_cfg4.txt 2> gpurun_out/r02_timeline2_cfg4.err; echo rc=$?; cat gpurun_out/r02_timeline2_cfg4.txt
