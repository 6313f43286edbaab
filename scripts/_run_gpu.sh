python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "== fast dense"; timeout 300 python scripts/dev_hope.py --n 1000000 --tol 1e-3 --reps 3 2>&1 | grep -o '"iters": [0-9]*\|"spmm_ms": [0-9.]*\|"dense_ms": [0-9.]*\|"spmm_count": [0-9]*\|"total_ms": [0-9.]*\|"resid_max": [0-9.e-]*' | tr '\n' ' '; echo
echo "== generic dense"; GEMB_DENSE_GENERIC=1 timeout 300 python scripts/dev_hope.py --n 1000000 --tol 1e-3 --reps 3 2>&1 | grep -o '"iters": [0-9]*\|"spmm_ms": [0-9.]*\|"dense_ms": [0-9.]*\|"spmm_count": [0-9]*\|"total_ms": [0-9.]*\|"resid_max": [0-9.e-]*' | tr '\n' ' '; echo
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_hope_r1w.json 2> gpurun_out/bench_hope_r1w.err; python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/bench_hope_r1w.json') if l.startswith('{')][-1])
print('value',j['value'],'ms',j['ms_per_step'],'e2e',j['e2e'])
PY
