#!/bin/bash
# third GPU call: ncu captures of the two headline kernels (cfg5 HELL, cfg2 HDIA)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD5="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
CMD2="python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD5 > gpurun_out/plain5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hell_spmv_kernel -s 3 -c 1 -f -o gpurun_out/prof_cfg5_r1 $CMD5 > gpurun_out/ncu5.log 2>&1
echo "ncu cfg5 rc=$?"; tail -3 gpurun_out/ncu5.log
$CMD2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hdia_spmv_kernel -s 3 -c 1 -f -o gpurun_out/prof_cfg2_r1 $CMD2 > gpurun_out/ncu2.log 2>&1
echo "ncu cfg2 rc=$?"; tail -3 gpurun_out/ncu2.log
$CMD5 > gpurun_out/plain5b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg5_r1.csv $CMD5 > gpurun_out/ncu5l.log 2>&1
echo "launch list rc=$?"; tail -12 gpurun_out/launches_cfg5_r1.csv
ls -la gpurun_out/
