mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_host_shim_gpu.py tests/test_frontend.py -m gpu -x -q > gpurun_out/c21_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/c21_pytest.log
