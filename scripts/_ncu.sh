# final round-1 evidence run (one GPU): default bench line, launch list, full capture of the dominant kernel
python bench.py > gpurun_out/bench_default_v4.log 2>gpurun_out/bench_default_v4.err || exit 1
tail -1 gpurun_out/bench_default_v4.log | cut -c1-600
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-krr --e2e-steps 1 > gpurun_out/bench_plain_v4.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_c2_v4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-krr > gpurun_out/ncu_launches_v4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kmm_tc_kernel -c 1 -o gpurun_out/r01_tc_c2_v4 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-krr --e2e-steps 1 > gpurun_out/ncu_full_v4.log 2>&1
ls -la gpurun_out | grep v4
