#!/bin/bash
# The GPU suite + smoke at HEAD -> gpurun_out/r02_pytest_gpu.log, r02_smoke.log (tools/summarize_profiles_r2.py copies them to profiles/)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
