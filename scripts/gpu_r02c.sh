mkdir -p gpurun_out
timeout 300 python scripts/exp_solver.py > gpurun_out/r02c_exp_solver.log 2>&1; echo "exp_solver rc=$?"
for v in v1 v3; do GEMB_SPMM=$v timeout 120 python scripts/exp_spmm.py >> gpurun_out/r02c_exp_spmm.log 2>&1; done
GEMB_SPMM=v3 GEMB_SPMM_PASSES=8 timeout 120 python scripts/exp_spmm.py >> gpurun_out/r02c_exp_spmm.log 2>&1
GEMB_SPMM=v3 GEMB_SPMM_PASSES=2 timeout 120 python scripts/exp_spmm.py >> gpurun_out/r02c_exp_spmm.log 2>&1
EXP_OS=8 GEMB_SPMM=v1 timeout 120 python scripts/exp_spmm.py >> gpurun_out/r02c_exp_spmm.log 2>&1
cat gpurun_out/r02c_exp_spmm.log
timeout 300 python scripts/dev_rmat.py --scale 20 > gpurun_out/r02c_rmat20.log 2>&1; echo "rmat rc=$?"; tail -12 gpurun_out/r02c_rmat20.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 1 --warmup 1 --no-node2vec --no-cpu --no-e2e > gpurun_out/r02c_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmm_bulk_kernel -s 30 -c 1 -o gpurun_out/r02c_spmm_bulk python bench.py --steps 1 --warmup 1 --no-node2vec --no-cpu --no-e2e > gpurun_out/r02c_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/r02c_exp_solver.log
