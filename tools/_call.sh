mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_descriptor_gpu.py -m gpu -x -q -s > gpurun_out/c3_pytest.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/c3_pytest.log
