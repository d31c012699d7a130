#!/bin/bash
# Profiling pass after the decoupled-ring GEMM (outputs under gpurun_out/, small).
set -x
O=gpurun_out
mkdir -p $O
M2="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 400 ncu --metrics $M2 --clock-control none -s 600 -c 500 --csv --log-file $O/r1c_bench_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/r1c_bench_under_ncu.log 2>&1
timeout 300 ncu --metrics $M2 --clock-control none --csv --log-file $O/r1c_potrf8192_launches.csv \
    python scripts/ncu_targets.py potrf 8192 > $O/r1c_potrf.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 4 -f -o /tmp/r1c_gemm4096 \
    python scripts/ncu_targets.py gemm > $O/r1c_gemm.log 2>&1
ncu -i /tmp/r1c_gemm4096.ncu-rep --page raw --csv > $O/r1c_gemm4096_raw.csv 2>/dev/null
ncu -i /tmp/r1c_gemm4096.ncu-rep --page source --csv --kernel-name regex:gemm_tc > $O/r1c_gemm4096_source.csv 2>/dev/null
timeout 300 ncu --metrics $M2 --clock-control none -k regex:"gemm_tc|kbuild|skinny" -c 60 --csv --log-file $O/r1c_sparsegp_launches.csv \
    python scripts/bench_sparsegp.py 56784 1024 --no-cpu > $O/r1c_sparsegp_ncu.log 2>&1
timeout 300 ncu --metrics $M2 --clock-control none -k regex:"mlp_tanh|normal_" -c 40 --csv --log-file $O/r1c_bnn_launches.csv \
    python scripts/bench_bnn.py 5 --no-cpu > $O/r1c_bnn_ncu.log 2>&1
du -sh $O
