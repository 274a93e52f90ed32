#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/timeline_dump.py --workload cfg2 --set "overlap=1" --set "overlap=0" > gpurun_out/r02_timeline_cfg2.txt 2> gpurun_out/r02_timeline_cfg2.err; echo rc=$?; cat gpurun_out/r02_timeline_cfg2.txt; tail -3 gpurun_out/r02_timeline_cfg2.err
