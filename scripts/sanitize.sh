#!/bin/bash
# compute-sanitizer passes over the hand-written kernels (run on the GPU box): memcheck (global / shared out-of-bounds,
# misaligned accesses) and racecheck (shared-memory hazards) on small instances of each kernel family + one captured step.
# Logs -> gpurun_out/sanitizer/ ; the summaries are committed under profiles/r2_sanitizer/.
set -x
O=gpurun_out/sanitizer
mkdir -p $O
export MXF_DAG_TIMEOUT_S=600
CS="compute-sanitizer --print-limit 20 --error-exitcode 9 --report-api-errors no"
run() {  # name, tool, pytest -k expression, test file
    timeout 900 $CS --tool $2 python -m pytest $4 -x -q -m gpu -k "$3" > $O/$1_$2.log 2>&1
    echo "$1 $2 rc=$?" >> $O/summary.txt
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/$1_$2.log | tail -3 >> $O/summary.txt
}
: > $O/summary.txt
run potrf_dag memcheck "potrf_dataflow_factor_and_inverse and (100 or 257 or 1024)" tests/test_gpu_kernels.py
run potrf_dag racecheck "potrf_dataflow_factor_and_inverse and (100 or 257)" tests/test_gpu_kernels.py
run potrf_packed memcheck "test_potrf_packed_and_pack_contents and f32 and (65 or 300 or 1536)" tests/test_gpu_kernels.py
run gemm_tc memcheck "test_trsm_solve_large_blocks and f32 and (256 or 768)" tests/test_gpu_kernels.py
run gemm_tc racecheck "test_trsm_solve_large_blocks and f32 and 256" tests/test_gpu_kernels.py
run kbuild memcheck "test_kbuild_cross and f32" tests/test_gpu_kernels.py
run kbuild racecheck "test_kbuild_cross and f32" tests/test_gpu_kernels.py
run mlp memcheck "mlp_tanh" tests/test_gpu_kernels.py
run mlp racecheck "mlp_tanh" tests/test_gpu_kernels.py
run step memcheck "test_svgp_minibatch_paths_agree" tests/test_gpu_api.py
cat $O/summary.txt
