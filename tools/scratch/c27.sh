mkdir -p gpurun_out/c27
timeout 600 python -m pytest tests/test_darknet_gpu.py -x -q -k "conv0" > gpurun_out/c27/tests.log 2>&1; echo "rc=$?" >> gpurun_out/c27/tests.log
tail -40 gpurun_out/c27/tests.log
timeout 300 python tools/conv0_cases.py > gpurun_out/c27/cases.log 2>&1; tail -14 gpurun_out/c27/cases.log
