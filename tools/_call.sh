mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c13_bench2.json 2> gpurun_out/c13_bench2.err; echo "bench2 exit $?"; python -c "
import json;b=json.load(open('gpurun_out/c13_bench2.json'));print(b['value'],b['ms_per_step'],b['e2e']['value'],b['stages_ms'],b['roofline']['kernel'],b['roofline']['frac'],b['gpu_launches'])"; tail -3 gpurun_out/c13_bench2.err
timeout 900 python -m pytest tests/test_sharding_cpu.py -x -q 2>&1 | tail -2
