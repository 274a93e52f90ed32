#!/bin/bash
# Round 2, pass 26: K2T row selection (dense_compact.cuh): parity tests, then A/B on cfg4 / cfg5 shard / cfg3-b256
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -k "row_selection or query_tiled" > gpurun_out/r02sel_pytest.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r02sel_pytest.log
for W in cfg4 cfg5-shard; do
timeout 600 python tools/ab_opts.py --workload $W --batches 6 --out gpurun_out/r02sel_ab_$W.jsonl --base "overlap=1" \
  --set "dense_compact=0" --set "dense_compact=70" --set "dense_compact=0" --set "dense_compact=70" 2> gpurun_out/r02sel_ab_$W.err | cut -c1-100,330-800; echo "$W rc=$?"; tail -2 gpurun_out/r02sel_ab_$W.err
done
