mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c22_bench4.json 2> gpurun_out/c22_bench4.err; echo "bench4 exit $?"; python -c "
import json
for l in open('gpurun_out/c22_bench4.json'):
    if l.startswith('{'):
        b=json.loads(l);print(b['value'],b['ms_per_step'],b['e2e']['value'],b['stages_ms'],b['roofline']['kernel'],b['roofline']['frac'],b['gpu_launches'])"; head -c 300 gpurun_out/c22_bench4.json; tail -3 gpurun_out/c22_bench4.err
