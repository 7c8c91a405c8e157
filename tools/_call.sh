mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_search_gpu.py tests/test_pipeline_gpu.py tests/test_state_io.py -m gpu -x -q > gpurun_out/c10_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/c10_pytest.log
timeout 300 python tools/bench_sweep_ab.py 2>&1 | grep -v V1 > gpurun_out/c10_ab.log; cat gpurun_out/c10_ab.log
CB_TOPK_ONE_PASS=1 timeout 300 python tools/bench_sweep_ab.py 2>&1 | grep -v V1 | head -2
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err; python -c "
import json;b=json.load(open('gpurun_out/c10_bench.json'));print(b['value'],b['e2e']['value'],b['stages_ms'],b['roofline']['frac'],b['roofline']['traffic'])"
