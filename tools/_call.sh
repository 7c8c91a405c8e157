mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_state_io.py tests/test_pipeline_gpu.py -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/c4_pytest.log
