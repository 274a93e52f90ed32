#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/timeline_dump.py --workload cfg4 --set "dense_wait_sparse=0" --set "dense_wait_sparse=1" > gpurun_out/r02_timeline_cfg4.txt 2> gpurun_out/r02_timeline_cfg4.err; echo rc=$?; cat gpurun_out/r02_timeline_cfg4.txt; tail -3 gpurun_out/r02_timeline_cfg4.err
