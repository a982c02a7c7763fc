#!/bin/bash
# The GPU suite, smoke() and the default bench under the final default BatchNorm order (25).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > $O/fc4_gputests.log 2>&1
echo "pytest rc=$?" >> $O/fc4_gputests.log
tail -3 $O/fc4_gputests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/fc4_smoke.log 2>&1
echo "smoke rc=$?" >> $O/fc4_smoke.log
tail -3 $O/fc4_smoke.log
timeout 200 python bench.py > $O/fc4_bench_default.json 2> $O/fc4_bench_default.err
echo "bench rc=$?"
cut -c1-300 $O/fc4_bench_default.json
