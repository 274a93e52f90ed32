#!/bin/bash
# Round 2, pass 27: row selection on every large segment, copies on a side stream: parity, bounds-checked build, A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -k "row_selection or query_tiled" > gpurun_out/r02sel2_pytest.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02sel2_pytest.log
VB200_LIB=$PWD/voitta-rag_b200/libvoitta_b200_dbg.so timeout 600 python tools/sanitize_workload.py > gpurun_out/r02sel2_bounds.log 2>&1; echo "bounds-checked build rc=$?"; tail -4 gpurun_out/r02sel2_bounds.log
VB200_LIB=$PWD/voitta-rag_b200/libvoitta_b200_dbg.so timeout 600 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -k "row_selection" > gpurun_out/r02sel2_pytest_bounds.log 2>&1; echo "row selection test on the bounds-checked build rc=$?"; tail -3 gpurun_out/r02sel2_pytest_bounds.log
for W in cfg4 cfg5-shard; do
timeout 600 python tools/ab_opts.py --workload $W --batches 6 --out gpurun_out/r02sel2_ab_$W.jsonl --base "overlap=1" \
  --set "dense_compact=0" --set "dense_compact=70" --set "dense_compact=0" --set "dense_compact=70" 2> gpurun_out/r02sel2_ab_$W.err | cut -c1-100,330-800; echo "$W rc=$?"; tail -2 gpurun_out/r02sel2_ab_$W.err
done
