mkdir -p gpurun_out
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
F='"ms_per_step": [0-9.]*|"phases_ms_per_step": {[^}]*}|"iters": [0-9]*|"resid_max_rel_sigma_max": [0-9.e-]*|"e2e": {"value": [0-9.]*, "unit": "nodes/s", "ms_per_step": [0-9.]*'
GEMB_WIRE=fp16 timeout 300 $T8 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-node2vec --no-e2e > gpurun_out/r02m_bench8_fp16.json 2> gpurun_out/r02m_bench8_fp16.err; echo "8 fp16 rc=$?"; grep -oE "$F" gpurun_out/r02m_bench8_fp16.json | head -5
timeout 300 $T4 --master-port 29532 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu --no-node2vec > gpurun_out/r02m_bench4_fp32.json 2> gpurun_out/r02m_bench4_fp32.err; echo "4 fp32 rc=$?"; grep -oE "$F" gpurun_out/r02m_bench4_fp32.json | head -6
GEMB_WIRE=fp16 timeout 300 $T4 --master-port 29533 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu --no-node2vec --no-e2e > gpurun_out/r02m_bench4_fp16.json 2> gpurun_out/r02m_bench4_fp16.err; echo "4 fp16 rc=$?"; grep -oE "$F" gpurun_out/r02m_bench4_fp16.json | head -5
