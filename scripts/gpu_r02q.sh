mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02q_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02q_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02q_smoke.log
timeout 400 python bench.py > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?"; cut -c1-2600 gpurun_out/r02q_bench.json; tail -3 gpurun_out/r02q_bench.err
( time timeout 700 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02q_bench_reference.json 2> gpurun_out/r02q_bench_reference.err ) 2>&1 | tail -3; echo "ref rc=$?"; cut -c1-1500 gpurun_out/r02q_bench_reference.json; tail -3 gpurun_out/r02q_bench_reference.err
