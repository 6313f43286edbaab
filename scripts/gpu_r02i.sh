mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 scripts/mgpu_check.py > gpurun_out/r02i_mgpu_check.log 2>&1; echo "check rc=$?"; tail -1 gpurun_out/r02i_mgpu_check.log | cut -c1-1500
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02i_bench2.json 2> gpurun_out/r02i_bench2.err; echo "bench2 rc=$?"; cut -c1-3200 gpurun_out/r02i_bench2.json; tail -3 gpurun_out/r02i_bench2.err
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --graph rmat --scale 20 --steps 3 --warmup 1 --no-cpu > gpurun_out/r02i_bench2_rmat20.json 2> gpurun_out/r02i_bench2_rmat20.err; echo "bench2 rmat rc=$?"; cut -c1-2400 gpurun_out/r02i_bench2_rmat20.json; tail -3 gpurun_out/r02i_bench2_rmat20.err
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --workload node2vec --graph rmat --scale 18 --warmup 1 --no-cpu > gpurun_out/r02i_bench2_n2v_rmat18.json 2> gpurun_out/r02i_bench2_n2v_rmat18.err; echo "bench2 n2v rmat rc=$?"; cut -c1-1800 gpurun_out/r02i_bench2_n2v_rmat18.json; tail -3 gpurun_out/r02i_bench2_n2v_rmat18.err
