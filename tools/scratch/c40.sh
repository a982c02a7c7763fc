mkdir -p gpurun_out/c40
timeout 120 python tools/gemm_cases.py 3 2>&1 | grep "L3"
for h in 0 1; do
AVDN_CONV_HALO=$h timeout 300 python bench.py --no-secondary --no-library-bar --no-cpu-baseline > gpurun_out/c40/bench_halo$h.json 2> gpurun_out/c40/bench_halo$h.err
python -c "
import json
d=json.loads(open('gpurun_out/c40/bench_halo$h.json').read()); b=d['roofline']['kernel_ms_breakdown']; print('halo=$h', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], {k:v['ms'] for k,v in list(b.items())[:6]}, d['gpu_launches']/d['steps'])"
done
