python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_hope_r1ac.json 2> gpurun_out/bench_hope_r1ac.err; tail -c 600 gpurun_out/bench_hope_r1ac.json
timeout 900 python bench.py --workload node2vec --steps 2 --warmup 3 > gpurun_out/bench_n2v_r1ac.json 2> gpurun_out/bench_n2v_r1ac.err; tail -c 400 gpurun_out/bench_n2v_r1ac.json
