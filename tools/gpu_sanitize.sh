#!/bin/bash
# compute-sanitizer is CLOSED on this GPU pool ("runs under it have left GPUs needing a reset"; the wrapper answers
# rc 86 — kept in gpurun_out/rNN_sanitize_version.log).  Substitute, as the pool's message suggests: bounds checks and
# asserts of our own.  libvoitta_b200_dbg.so is the library built with -DVB_DEBUG_BOUNDS (voitta-rag_b200/build.py
# build_debug()): every VB_CHECK in the kernels is a device-side assert.  The all-kernel workload runs against it and
# against the release build; both must finish with every result equal to the oracle's.
R=${1:-r02}
mkdir -p gpurun_out
compute-sanitizer --version > gpurun_out/${R}_sanitize_version.log 2>&1; echo "compute-sanitizer rc=$?"; head -2 gpurun_out/${R}_sanitize_version.log
timeout 600 python tools/sanitize_workload.py > gpurun_out/${R}_sanitize_plain.log 2>&1; echo "release build rc=$?"; tail -2 gpurun_out/${R}_sanitize_plain.log
VB200_LIB=$PWD/voitta-rag_b200/libvoitta_b200_dbg.so timeout 900 python tools/sanitize_workload.py > gpurun_out/${R}_sanitize_bounds.log 2>&1; echo "bounds-checked build rc=$?"; tail -3 gpurun_out/${R}_sanitize_bounds.log
VB200_LIB=$PWD/voitta-rag_b200/libvoitta_b200_dbg.so timeout 1500 python -m pytest tests -m gpu -q -x -k "not full_size and not model_dimensions" > gpurun_out/${R}_pytest_bounds.log 2>&1; echo "GPU suite on the bounds-checked build rc=$?"; tail -3 gpurun_out/${R}_pytest_bounds.log
