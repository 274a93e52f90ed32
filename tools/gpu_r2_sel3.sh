#!/bin/bash
# Round 2, pass 28: row selection in the resident K2 kernel too, auto threshold: parity, then bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -k "row_selection or query_tiled or dimensions" > gpurun_out/r02sel3_pytest.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02sel3_pytest.log
for W in cfg3-b256-s1 cfg3-b256-s50; do
timeout 600 python tools/ab_opts.py --workload $W --batches 6 --out gpurun_out/r02sel3_ab_$W.jsonl --base "overlap=1" \
  --set "dense_compact=0" --set "dense_compact=-1" --set "dense_compact=0" --set "dense_compact=-1" 2> gpurun_out/r02sel3_ab_$W.err | cut -c1-100,330-800; echo "$W rc=$?"; tail -2 gpurun_out/r02sel3_ab_$W.err
done
timeout 600 python tools/ab_opts.py --workload cfg4 --batches 6 --out gpurun_out/r02sel3_ab_cfg4.jsonl --base "overlap=1" \
  --set "dense_compact=-1" --set "dense_compact=0" 2> gpurun_out/r02sel3_ab_cfg4.err | cut -c1-100,330-800; echo "cfg4 rc=$?"; tail -2 gpurun_out/r02sel3_ab_cfg4.err
