#!/bin/bash
# Final profiling pass of round 1 (r1d): current kernels.  Outputs under gpurun_out/ (small).
set -x
O=gpurun_out
mkdir -p $O
M2="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 400 ncu --metrics $M2 --clock-control none -s 700 -c 420 --csv --log-file $O/r1d_bench_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/r1d_bench_under_ncu.log 2>&1
timeout 300 ncu --metrics $M2 --clock-control none --csv --log-file $O/r1d_potrf8192_launches.csv \
    python scripts/ncu_targets.py potrf 8192 > $O/r1d_potrf.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pe -c 3 -f -o /tmp/r1d_gemm_pe \
    python scripts/microbench.py gemmshapes > $O/r1d_gemm_pe.log 2>&1
ncu -i /tmp/r1d_gemm_pe.ncu-rep --page raw --csv > $O/r1d_gemm_pe_raw.csv 2>/dev/null
timeout 300 ncu --metrics $M2 --clock-control none -k regex:"gemm_tc|kbuild|skinny" -c 60 --csv --log-file $O/r1d_sparsegp_launches.csv \
    python scripts/bench_sparsegp.py 56784 1024 --no-cpu > $O/r1d_sparsegp_ncu.log 2>&1
timeout 300 ncu --metrics $M2 --clock-control none -k regex:"mlp_tanh|normal_|mt_" -c 40 --csv --log-file $O/r1d_bnn_launches.csv \
    python scripts/bench_bnn.py 5 --no-cpu > $O/r1d_bnn_ncu.log 2>&1
du -sh $O
