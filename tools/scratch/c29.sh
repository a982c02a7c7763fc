mkdir -p gpurun_out/c29
timeout 300 python tools/conv0_cases.py > gpurun_out/c29/cases.log 2>&1; tail -6 gpurun_out/c29/cases.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c29/tests.log 2>&1; echo "rc=$?" >> gpurun_out/c29/tests.log
tail -15 gpurun_out/c29/tests.log
timeout 600 python bench.py --no-secondary --no-library-bar --no-cpu-baseline > gpurun_out/c29/bench.json 2> gpurun_out/c29/bench.err
python -c "
import json
d=json.loads(open('gpurun_out/c29/bench.json').read()); print(round(d['ms_per_step'],2), round(d['value'],1), d['clocks']); print(d['roofline']['kernel_ms_breakdown'])"
