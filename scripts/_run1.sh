timeout 300 python scripts/tc_diag.py basic > gpurun_out/v3_basic.log 2>&1; echo "basic exit $?"
tail -16 gpurun_out/v3_basic.log
timeout 300 python scripts/tc_diag.py range > gpurun_out/v3_range.log 2>&1; echo "range exit $?"
tail -10 gpurun_out/v3_range.log
timeout 300 python scripts/tc_diag.py perf > gpurun_out/v3_perf.log 2>&1; echo "perf exit $?"
cat gpurun_out/v3_perf.log
timeout 300 python scripts/tc_diag.py knock > gpurun_out/v3_knock.log 2>&1; echo "knock exit $?"
cat gpurun_out/v3_knock.log
