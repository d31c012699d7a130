#!/bin/bash
# Round-2 (second half) profiling pass on one GPU: launch list of the headline bench + full captures of the tcgen05 K-build.
set -x
O=gpurun_out
mkdir -p $O
M2="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 400 ncu --metrics $M2 --clock-control none -c 4000 --csv --log-file $O/r2b_bench_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2b_bench_under_ncu.log 2>&1
# K(X,Z) N=1e6 M=1024 on the tensor-core kernel: the headline shape (D=8 RBF) and config 3's (D=16 Matern-5/2)
timeout 200 ncu --set full --clock-control none --import-source on -k regex:kbuild_fwd_tc -c 1 -s 2 -f -o /tmp/kt8 \
    python scripts/ncu_targets.py kbuild_tc 8 0 > $O/r2b_kbuild_tc_rbf8.log 2>&1
ncu -i /tmp/kt8.ncu-rep --page raw --csv > $O/r2b_kbuild_tc_rbf8_raw.csv 2>/dev/null
timeout 200 ncu --set full --clock-control none --import-source on -k regex:kbuild_fwd_tc -c 1 -s 2 -f -o /tmp/kt16 \
    python scripts/ncu_targets.py kbuild_tc 16 3 > $O/r2b_kbuild_tc_m52_16.log 2>&1
ncu -i /tmp/kt16.ncu-rep --page raw --csv > $O/r2b_kbuild_tc_m52_16_raw.csv 2>/dev/null
cuobjdump -sass mxfusion_b200/libmxf_b200.so 2>/dev/null | grep -E "Function : .*kbuild_fwd_tc|UTCHMMA|UTMASTG|LDTM|UTMALDG" | awk '/Function/ {f=$0} !/Function/ {c[f" "$2]++} END {for (k in c) print c[k], k}' | sort -k2 > $O/r2b_sass_mnemonics.txt
du -sh $O
