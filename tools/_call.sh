mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_descriptor_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > gpurun_out/c11_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/c11_pytest.log
timeout 300 python tools/bench_desc.py
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv1 -c 3 python tools/bench_desc.py 2>&1 | grep -E "gpu__time" | head -5
