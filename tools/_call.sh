mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_frontend.py tests/test_abi.py -x -q > gpurun_out/c6_pytest.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/c6_pytest.log
