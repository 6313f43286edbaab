timeout 900 python -m pytest tests/test_gpu_hope.py tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python scripts/dev_hope.py --n 1000000 --tol 1e-3 --reps 3 2>&1 | grep -o '"iters": [0-9]*\|"spmm_ms": [0-9.]*\|"dense_ms": [0-9.]*\|"spmm_count": [0-9]*\|"total_ms": [0-9.]*\|"resid_max": [0-9.e-]*' | tr '\n' ' '; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_hope_r1z.csv python scripts/dev_hope.py --n 1000000 --tol 1e-3 --reps 1 > gpurun_out/ncu_r1z.log 2>&1
