mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -x -q > gpurun_out/c12_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/c12_pytest.log
timeout 300 python tools/bench_sweep_ab.py > gpurun_out/c12_ab.log 2>&1; cat gpurun_out/c12_ab.log
