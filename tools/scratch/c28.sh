mkdir -p gpurun_out/c28
timeout 600 python -m pytest tests/test_darknet_gpu.py -q -k "conv0" > gpurun_out/c28/tests.log 2>&1; echo "rc=$?" >> gpurun_out/c28/tests.log
tail -40 gpurun_out/c28/tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv0_tc_ --launch-skip 5 --launch-count 5 -o gpurun_out/c28/conv0_tc -f python tools/conv0_cases.py > gpurun_out/c28/ncu.log 2>&1; tail -3 gpurun_out/c28/ncu.log
