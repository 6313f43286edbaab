mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lap.py -x -q > gpurun_out/r02n_pytest_lap.log 2>&1; echo "lap rc=$?"; tail -25 gpurun_out/r02n_pytest_lap.log
