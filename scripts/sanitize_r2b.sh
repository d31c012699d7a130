#!/bin/bash
# compute-sanitizer on the kernels added in the second half of round 2: the tcgen05 + TMA-store K-build (cross and symmetric
# paths) and the step that now runs it.  Logs -> gpurun_out/sanitizer_r2b/ ; summary committed under profiles/r2_sanitizer/.
set -x
O=gpurun_out/sanitizer_r2b
mkdir -p $O
CS="compute-sanitizer --print-limit 20 --error-exitcode 9 --report-api-errors no"
run() {  # name, tool, pytest -k expression, test file
    timeout 600 $CS --tool $2 python -m pytest $4 -x -q -m gpu -k "$3" > $O/$1_$2.log 2>&1
    echo "$1 $2 rc=$?" >> $O/summary.txt
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/$1_$2.log | tail -3 >> $O/summary.txt
}
: > $O/summary.txt
run kbuild_tc memcheck "tensor_core_path and not shape5" tests/test_gpu_kernels.py
run kbuild_tc racecheck "tensor_core_path and (shape0 or shape1 or shape2) and not shape5" tests/test_gpu_kernels.py
run step_tc memcheck "test_svgp_minibatch_paths_agree" tests/test_gpu_api.py
cat $O/summary.txt
