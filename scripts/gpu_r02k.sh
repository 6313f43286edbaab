mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r02k_topo.txt 2>&1
timeout 420 $TR --master-port 29521 bench.py --gpus 8 --graph rmat --scale 24 --steps 3 --warmup 1 --no-cpu > gpurun_out/r02k_bench8_hope_rmat24.json 2> gpurun_out/r02k_bench8_hope_rmat24.err; echo "hope rmat24 rc=$?"; cut -c1-2600 gpurun_out/r02k_bench8_hope_rmat24.json; tail -4 gpurun_out/r02k_bench8_hope_rmat24.err
timeout 420 $TR --master-port 29522 bench.py --gpus 8 --workload node2vec --graph rmat --scale 24 --warmup 1 --no-cpu > gpurun_out/r02k_bench8_n2v_rmat24.json 2> gpurun_out/r02k_bench8_n2v_rmat24.err; echo "n2v rmat24 rc=$?"; cut -c1-2200 gpurun_out/r02k_bench8_n2v_rmat24.json; tail -4 gpurun_out/r02k_bench8_n2v_rmat24.err
timeout 420 $TR --master-port 29523 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-node2vec > gpurun_out/r02k_bench8_sbm.json 2> gpurun_out/r02k_bench8_sbm.err; echo "sbm8 rc=$?"; cut -c1-3000 gpurun_out/r02k_bench8_sbm.json; tail -4 gpurun_out/r02k_bench8_sbm.err
