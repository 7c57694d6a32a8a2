#!/bin/bash
# ncu --set full of the HDIA variants and of plain DIA on the cfg2 matrix (1 GPU)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$B --workload cfg2 --tune hdiaVariant=5 > gpurun_out/plain_v5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hdia_spmv -s 3 -c 1 -f -o gpurun_out/prof_r1_cfg2_hdia_v5 $B --workload cfg2 --tune hdiaVariant=5 > gpurun_out/ncu_v5.log 2>&1; echo "v5 rc=$?"
$B --workload cfg2 --tune hdiaVariant=6 > gpurun_out/plain_v6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hdia_spmv -s 3 -c 1 -f -o gpurun_out/prof_r1_cfg2_hdia_v6 $B --workload cfg2 --tune hdiaVariant=6 > gpurun_out/ncu_v6.log 2>&1; echo "v6 rc=$?"
$B --workload cfg2dia > gpurun_out/plain_dia.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dia_spmv -s 3 -c 1 -f -o gpurun_out/prof_r1_cfg2_dia $B --workload cfg2dia > gpurun_out/ncu_dia.log 2>&1; echo "dia rc=$?"
$B --workload cfg2 > gpurun_out/plain_v0.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hdia_spmv -s 3 -c 1 -f -o gpurun_out/prof_r1_cfg2_hdia_v0 $B --workload cfg2 > gpurun_out/ncu_v0.log 2>&1; echo "v0 rc=$?"
ls -la gpurun_out/*.ncu-rep
