mkdir -p gpurun_out/c41
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3_halo --launch-skip 3 --launch-count 1 -o gpurun_out/c41/halo_fwd -f python tools/gemm_cases.py 3 > gpurun_out/c41/ncu1.log 2>&1; tail -2 gpurun_out/c41/ncu1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3_halo --launch-skip 7 --launch-count 1 -o gpurun_out/c41/halo_dgrad -f python tools/gemm_cases.py 3 > gpurun_out/c41/ncu2.log 2>&1; tail -2 gpurun_out/c41/ncu2.log
