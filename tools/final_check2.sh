#!/bin/bash
# Second one-call check: the GPU suite and smoke() under the new default BatchNorm order, a longer A/B between the
# reversed orders, and the default bench under the winner (9 unless another mask beats it in every round).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > $O/fc2_gputests.log 2>&1
echo "pytest rc=$?" >> $O/fc2_gputests.log
tail -3 $O/fc2_gputests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/fc2_smoke.log 2>&1
echo "smoke rc=$?" >> $O/fc2_smoke.log
tail -3 $O/fc2_smoke.log
timeout 100 python tools/bn_order_ab.py 12 3 9,13,15,11,0 > $O/fc2_bn_order_ab.txt 2> $O/fc2_bn_order_ab.err
echo "ab rc=$?"
BEST=$(python - <<'EOF'
import json
try:
    d = json.loads(open("gpurun_out/fc2_bn_order_ab.txt").read().strip().splitlines()[-1])["median_best_ms"]
    base = [r[0] for r in d["9"]]
    best, gain = 9, 0.0
    for m, rs in d.items():
        if m in ("9", "0"):
            continue
        med = [r[0] for r in rs]
        if all(x < b for x, b in zip(med, base)):
            g = 1.0 - sum(med) / sum(base)
            if g > 0.005 and g > gain:
                best, gain = int(m), g
    print(best)
except Exception:
    print(9)
EOF
)
echo "chosen AVDN_BN_ORDER=$BEST" | tee $O/fc2_bn_order_choice.txt
AVDN_BN_ORDER=$BEST timeout 200 python bench.py > $O/fc2_bench_default.json 2> $O/fc2_bench_default.err
echo "bench rc=$?"
cut -c1-300 $O/fc2_bench_default.json
