#!/bin/bash
# after a default change: full gpu test-suite, cfg1/cfg2 bench lines, ncu captures of both kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for c in cfg2 cfg1; do
  timeout 600 python bench.py --workload $c --steps 20 --warmup 3 > gpurun_out/bench_r1_${c}_n1.json 2> gpurun_out/bench_r1_${c}_n1.err; echo "$c rc=$?"; cat gpurun_out/bench_r1_${c}_n1.json
done
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$B --workload cfg2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hdia_spmv -s 3 -c 1 -f -o gpurun_out/prof_r1_cfg2_hdia $B --workload cfg2 > gpurun_out/ncu2.log 2>&1; echo "ncu cfg2 rc=$?"
$B --workload cfg1 > gpurun_out/plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ell_spmv -s 3 -c 1 -f -o gpurun_out/prof_r1_cfg1_ell $B --workload cfg1 > gpurun_out/ncu1.log 2>&1; echo "ncu cfg1 rc=$?"
timeout 300 python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu --no-e2e --sweep "ellRows=-1;ellRows=0;ellRows=2;ellRows=0" > /dev/null 2> gpurun_out/sweep_cfg1.err; grep sweep gpurun_out/sweep_cfg1.err
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu --no-e2e --sweep "hdiaBlock=8;hdiaBlock=0;hdiaBlock=176;hdiaBlock=0" > /dev/null 2> gpurun_out/sweep_cfg2.err; grep sweep gpurun_out/sweep_cfg2.err
