for W in c3_matern52 c3_laplace c5_sketch; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 8 --steps 2 --warmup 1 --e2e-steps 1 --workload $W 2>&1 | tail -1 > gpurun_out/bench_${W}_n8.log
  cut -c1-260 gpurun_out/bench_${W}_n8.log
done
