#!/bin/bash
# One GPU-box call that (1) runs the GPU suite, (2) A/Bs the BatchNorm traversal orders inside the training step,
# (3) runs the default bench under the order that won (AVDN_BN_ORDER, 0 unless a mask is faster in both rounds),
# (4) takes the ncu launch list of one step.  Every stage has its own limit; outputs land in gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/fc_gpu.txt 2>&1
timeout 240 python -m pytest tests -m gpu -x -q -k "not traversal_orders" > $O/fc_gputests.log 2>&1
echo "pytest rc=$?" >> $O/fc_gputests.log
tail -3 $O/fc_gputests.log
timeout 90 python -m pytest tests/test_darknet_gpu.py -m gpu -q -k "traversal_orders" > $O/fc_gputests_orders.log 2>&1
echo "pytest rc=$?" >> $O/fc_gputests_orders.log
tail -15 $O/fc_gputests_orders.log
timeout 120 python tools/bn_order_ab.py 8 > $O/fc_bn_order_ab.txt 2> $O/fc_bn_order_ab.err
echo "ab rc=$?"
BEST=$(python - <<'EOF'
import json
try:
    d = json.loads(open("gpurun_out/fc_bn_order_ab.txt").read().strip().splitlines()[-1])["median_best_ms"]
    base = [r[0] for r in d["0"]]
    best, gain = 0, 0.0
    for m, rs in d.items():
        if m == "0":
            continue
        med = [r[0] for r in rs]
        # faster than the default order in BOTH rounds, by more than 0.4 % on average
        if all(x < b for x, b in zip(med, base)):
            g = 1.0 - sum(med) / sum(base)
            if g > 0.004 and g > gain:
                best, gain = int(m), g
    print(best)
except Exception:
    print(0)
EOF
)
echo "chosen AVDN_BN_ORDER=$BEST" | tee $O/fc_bn_order_choice.txt
AVDN_BN_ORDER=$BEST timeout 300 python bench.py > $O/fc_bench_default.json 2> $O/fc_bench_default.err
echo "bench rc=$?"
cat $O/fc_bench_default.json | cut -c1-600
AVDN_BN_ORDER=$BEST timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 --launch-count 1400 --csv \
  --log-file $O/fc_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-library-bar \
  > $O/fc_ncu_stdout.log 2>&1
echo "ncu rc=$?"
