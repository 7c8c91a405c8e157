mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_frontend.py -m gpu -x -q > gpurun_out/c17_pytest.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/c17_pytest.log
