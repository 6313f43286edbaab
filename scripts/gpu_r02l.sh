mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29511 scripts/mgpu_check.py > gpurun_out/r02l_mgpu_check.log 2>&1; echo "check rc=$?"; tail -1 gpurun_out/r02l_mgpu_check.log | cut -c1-2500
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-node2vec > gpurun_out/r02l_bench2.json 2> gpurun_out/r02l_bench2.err; echo "bench2 rc=$?"; cut -c1-3300 gpurun_out/r02l_bench2.json; tail -3 gpurun_out/r02l_bench2.err
GEMB_WIRE=fp32 timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-node2vec --no-e2e > gpurun_out/r02l_bench2_fp32wire.json 2> gpurun_out/r02l_bench2_fp32wire.err; echo "bench2 fp32 rc=$?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02l_bench2_fp32wire.json | head -1; grep -o '"phases_ms_per_step": {[^}]*}' gpurun_out/r02l_bench2_fp32wire.json
