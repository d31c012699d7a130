#!/bin/bash
# Round-1 profiling pass (run under gpurun, one GPU).  Outputs under gpurun_out/ (kept small: the raw / source pages
# are exported to CSV on the box and the large .ncu-rep files are dropped).
set -x
O=gpurun_out
mkdir -p $O
M2="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum"
# 1. launch list of the bench command (time + tensor-pipe activity + DRAM bytes per launch)
timeout 600 ncu --metrics $M2 --clock-control none -s 600 -c 500 --csv --log-file $O/r1b_bench_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/r1b_bench_under_ncu.log 2>&1
# 2. potrf N=8192 launch list
timeout 600 ncu --metrics $M2 --clock-control none --csv --log-file $O/r1b_potrf8192_launches.csv \
    python scripts/ncu_targets.py potrf 8192 > $O/r1b_potrf.log 2>&1
# 3. full captures: TMEM-A GEMM at 4096^3, K-build
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 4 -f -o /tmp/r1b_gemm4096 \
    python scripts/ncu_targets.py gemm > $O/r1b_gemm.log 2>&1
ncu -i /tmp/r1b_gemm4096.ncu-rep --page raw --csv > $O/r1b_gemm4096_raw.csv 2>/dev/null
ncu -i /tmp/r1b_gemm4096.ncu-rep --page source --csv --kernel-name regex:gemm_tc > $O/r1b_gemm4096_source.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:kbuild_fwd -c 3 -f -o /tmp/r1b_kbuild \
    python scripts/ncu_targets.py kbuild > $O/r1b_kbuild.log 2>&1
ncu -i /tmp/r1b_kbuild.ncu-rep --page raw --csv > $O/r1b_kbuild_raw.csv 2>/dev/null
ls -la $O /tmp/*.ncu-rep
du -sh $O
