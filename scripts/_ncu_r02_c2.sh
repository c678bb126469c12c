#!/bin/bash
# round-2 evidence for the headline workload (one GPU): plain run, launch list of the same command, full capture of the dominant kernel
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-krr --no-secondary --e2e-steps 1"
$CMD > gpurun_out/r02_bench_plain_c2.log 2> gpurun_out/r02_bench_plain_c2.err || exit 1
tail -c 700 gpurun_out/r02_bench_plain_c2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c2.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kmm_tc_kernel -s 1 -c 1 -f -o gpurun_out/r02_tc_c2 $CMD > gpurun_out/r02_ncu_full_c2.log 2>&1
ls -la gpurun_out/r02_tc_c2.ncu-rep gpurun_out/r02_launches_c2.csv
