#!/bin/bash
# usage: scripts/_scale_r02.sh N  -- the bench lines of round 2 at N GPUs (run under gpurun --gpus N)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_n2.log
fi
if [ "$N" = "8" ]; then
  $TR --master-port 29524 scripts/_e2e_breakdown.py > gpurun_out/r02_e2e_breakdown_n8.log 2>&1; echo "breakdown rc=$?"; cat gpurun_out/r02_e2e_breakdown_n8.log | tail -14
  $TR --master-port 29520 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_c2_n$N.json 2> gpurun_out/r02_bench_c2_n$N.err; echo "c2 rc=$?"
  $TR --master-port 29523 bench.py --gpus $N --steps 2 --warmup 1 --workload c5_sketch --no-krr > gpurun_out/r02_bench_c5_sketch_n$N.json 2> gpurun_out/r02_bench_c5_sketch_n$N.err; echo "c5 rc=$?"
fi
$TR --master-port 29521 bench.py --gpus $N --steps 2 --warmup 1 --workload c3_laplace --no-krr > gpurun_out/r02_bench_c3_laplace_n$N.json 2> gpurun_out/r02_bench_c3_laplace_n$N.err; echo "laplace rc=$?"
$TR --master-port 29522 bench.py --gpus $N --steps 3 --warmup 2 --workload c3_matern52 --no-krr > gpurun_out/r02_bench_c3_matern52_n$N.json 2> gpurun_out/r02_bench_c3_matern52_n$N.err; echo "matern rc=$?"
for f in gpurun_out/r02_bench_*_n$N.json; do echo "== $f"; tail -c 1500 $f; echo; done
for f in gpurun_out/r02_bench_*_n$N.err; do tail -n 2 $f; done
