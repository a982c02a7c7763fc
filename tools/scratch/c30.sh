mkdir -p gpurun_out/c30
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/c30/tests.log 2>&1; echo "rc=$?" >> gpurun_out/c30/tests.log
tail -25 gpurun_out/c30/tests.log
