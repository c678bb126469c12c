python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-krr --e2e-steps 1 > gpurun_out/bench_plain_v3b.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_c2_v3b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-krr > gpurun_out/ncu_launches_v3b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kmm_tc_kernel -c 1 -o gpurun_out/r01_tc_c2_v3b -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-krr --e2e-steps 1 > gpurun_out/ncu_full_v3b.log 2>&1
ls -la gpurun_out | tail -5
