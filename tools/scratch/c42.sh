timeout 200 python -m pytest tests/test_gemm_gpu.py -x -q -k "thin_halo" 2>&1 | tail -8
timeout 120 python tools/gemm_cases.py 3 2>&1 | grep "L3"
