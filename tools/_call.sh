mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_descriptor_gpu.py -m gpu -x -q > gpurun_out/c8_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/c8_pytest.log
timeout 300 python tools/bench_desc.py > gpurun_out/c8_desc.log 2>&1; cat gpurun_out/c8_desc.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vlad -c 12 python tools/bench_desc.py 2>&1 | grep -E "vlad|gpu__time" | head -30
