#!/bin/bash
# round-2 ncu evidence (one GPU): plain runs first, then one full capture per kernel family
set -x
T1="LaplaceLinOp 56832 262144 32 16"; T2="Matern52LinOp 37888 1000000 32 16"; T3="RBFLinOp 37888 2000000 16 1"; T4="RBFLinOp 18944 262144 64 1000"
for T in "$T1" "$T2" "$T3" "$T4"; do python scripts/_ncu_target.py $T || exit 1; done > gpurun_out/r02_ncu_plain.log 2>&1
cat gpurun_out/r02_ncu_plain.log
NCU="ncu --set full --clock-control none --import-source on -s 2 -c 1 -f"
$NCU -k regex:kmm_simt_kernel -o gpurun_out/r02_simt_laplace python scripts/_ncu_target.py $T1 > gpurun_out/r02_ncu_1.log 2>&1
$NCU -k regex:kmm_tc_kernel -o gpurun_out/r02_tc_matern52_c3 python scripts/_ncu_target.py $T2 > gpurun_out/r02_ncu_2.log 2>&1
$NCU -k regex:kmm_tc_kernel -o gpurun_out/r02_tc_kv_k1 python scripts/_ncu_target.py $T3 > gpurun_out/r02_ncu_3.log 2>&1
$NCU -k regex:kmm_tc_kernel -o gpurun_out/r02_tc_k1000 python scripts/_ncu_target.py $T4 > gpurun_out/r02_ncu_4.log 2>&1
tail -3 gpurun_out/r02_ncu_?.log; ls -la gpurun_out/*.ncu-rep
