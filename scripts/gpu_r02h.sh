mkdir -p gpurun_out
for v in v1 v3; do EXP_OS=8 GEMB_SPMM=$v timeout 120 python scripts/exp_spmm.py >> gpurun_out/r02h_exp_spmm.log 2>&1; done
EXP_OS=8 GEMB_SPMM=v3 GEMB_SPMM_PASSES=3 timeout 120 python scripts/exp_spmm.py >> gpurun_out/r02h_exp_spmm.log 2>&1
EXP_OS=8 GEMB_SPMM=v3 GEMB_SPMM_PASSES=6 timeout 120 python scripts/exp_spmm.py >> gpurun_out/r02h_exp_spmm.log 2>&1
cat gpurun_out/r02h_exp_spmm.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02h_rmat20_launches.csv python scripts/dev_rmat.py --scale 20 --reps 1 > gpurun_out/r02h_rmat20_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02h_rmat20_ncu.log
