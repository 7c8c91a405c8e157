mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_frontend.py -m gpu -x -q --durations=0 > gpurun_out/c19_pytest.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/c19_pytest.log
