#!/bin/bash
# The register-prefetching LSTM-side kernels (avdn_lstm_set_kernels): bit-level tests and the rollout bench.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 60 python -m pytest tests/test_lstm_gpu.py -m gpu -q > $O/fc5_gputests_lstm.log 2>&1
echo "pytest rc=$?" >> $O/fc5_gputests_lstm.log
tail -6 $O/fc5_gputests_lstm.log
timeout 50 python bench.py --workload rollout --no-cpu-baseline > $O/fc5_bench_rollout.json 2> $O/fc5_bench_rollout.err
echo "bench rc=$?"
cut -c1-250 $O/fc5_bench_rollout.json
