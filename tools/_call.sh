mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c15_bench.json 2> gpurun_out/c15_bench.err; tail -3 gpurun_out/c15_bench.err; python -c "
import json
for l in open('gpurun_out/c15_bench.json'):
    if l.startswith('{'):
        b=json.loads(l);print(b['value'],b['ms_per_step'],b['e2e'],b['stages_ms'],b['roofline']['frac'],b['gpu_launches'])"
