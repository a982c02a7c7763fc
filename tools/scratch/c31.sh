mkdir -p gpurun_out/c31
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/c31/tests.log 2>&1; echo "rc=$?" >> gpurun_out/c31/tests.log
tail -8 gpurun_out/c31/tests.log
for pdl in 0 1 0 1; do
AVDN_PDL=$pdl timeout 600 python bench.py --no-secondary --no-library-bar --no-cpu-baseline > gpurun_out/c31/bench_pdl$pdl.json 2> gpurun_out/c31/bench_pdl$pdl.err
python -c "
import json
d=json.loads(open('gpurun_out/c31/bench_pdl$pdl.json').read()); print('pdl=$pdl', round(d['ms_per_step'],2), round(d['value'],1), d['clocks']['sm_mhz']); b=d['roofline']['kernel_ms_breakdown']; print({k:v['ms'] for k,v in b.items()})"
done
