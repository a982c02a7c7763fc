#!/bin/bash
# Third one-call check: the unroll-8 variants of the BatchNorm passes (order bit 16): bit-level test, in-step A/B.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 90 python -m pytest tests/test_darknet_gpu.py -m gpu -q -x -k "traversal_orders" > $O/fc3_gputests_orders.log 2>&1
echo "pytest rc=$?" >> $O/fc3_gputests_orders.log
tail -4 $O/fc3_gputests_orders.log
timeout 100 python tools/bn_order_ab.py 12 3 9,25,27,31 > $O/fc3_bn_order_ab.txt 2> $O/fc3_bn_order_ab.err
echo "ab rc=$?"
grep -v "^{" $O/fc3_bn_order_ab.txt
