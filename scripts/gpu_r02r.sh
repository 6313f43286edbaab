mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02r_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02r_smoke.log
timeout 400 python bench.py --no-cpu > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; echo "bench rc=$?"
python - <<PY
import json
j=json.loads(open("gpurun_out/r02r_bench.json").read().strip().split("\n")[-1])
print({k:j[k] for k in ("value","ms_per_step","steps","clocks","gpu_launches")}, j["phases_ms_per_step"], j["e2e"]["ms_per_step"], j["e2e"]["min_ms_per_step"], j["roofline"]["frac"], j["roofline"]["ms_per_launch"], j["node2vec"]["value"], j["node2vec"]["roofline"]["traffic"])
PY
tail -2 gpurun_out/r02r_bench.err
