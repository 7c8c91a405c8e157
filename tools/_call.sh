mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_search_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/c2_pytest.log
timeout 300 python tools/bench_sweep_ab.py > gpurun_out/c2_sweep_ab.log 2>&1; cat gpurun_out/c2_sweep_ab.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; cat gpurun_out/c2_bench.json; tail -3 gpurun_out/c2_bench.err
