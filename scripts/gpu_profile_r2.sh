#!/bin/bash
# Round-2 profiling pass (one GPU).  Outputs under gpurun_out/ (small CSVs; the summaries are committed under profiles/).
set -x
O=gpurun_out
mkdir -p $O
M2="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum"
# every launch of the headline bench (3 timed steps after warm-up + the stand-alone roofline launches)
timeout 500 ncu --metrics $M2 --clock-control none -c 4000 --csv --log-file $O/r2_bench_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_under_ncu.log 2>&1
# potrf N = 8192: the dataflow kernel on the 1024-blocks + the tcgen05 panel / trailing GEMMs with their tensor-pipe activity
timeout 300 ncu --metrics $M2 --clock-control none --csv --log-file $O/r2_potrf8192_launches.csv \
    python scripts/ncu_targets.py potrf 8192 > $O/r2_potrf.log 2>&1
timeout 200 ncu --metrics $M2 --clock-control none --csv --log-file $O/r2_potrf1024_launches.csv \
    python scripts/ncu_targets.py potrf 1024 > $O/r2_potrf1024.log 2>&1
# full capture of the K(X,Z) stream kernel (DRAM traffic per launch) and of the dataflow potrf
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kbuild_fwd_stream -c 2 -f -o /tmp/r2_kbuild \
    python scripts/ncu_targets.py kbuild > $O/r2_kbuild.log 2>&1
ncu -i /tmp/r2_kbuild.ncu-rep --page raw --csv > $O/r2_kbuild_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:potrf_dag -c 1 -f -o /tmp/r2_potrf_dag \
    python scripts/ncu_targets.py potrf 1024 > $O/r2_potrf_dag.log 2>&1
ncu -i /tmp/r2_potrf_dag.ncu-rep --page raw --csv > $O/r2_potrf_dag_raw.csv 2>/dev/null
du -sh $O
