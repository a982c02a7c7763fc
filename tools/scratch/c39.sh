timeout 200 python -m pytest tests/test_gemm_gpu.py -x -q -k "thin_halo" 2>&1 | tail -15
timeout 400 python -m pytest tests/test_darknet_gpu.py tests/test_agent_gpu.py -x -q 2>&1 | tail -6
