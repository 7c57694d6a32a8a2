#!/bin/bash
# round-1 evidence run (1 GPU): full gpu test-suite, bench lines of all five configs, ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --cg > gpurun_out/bench_r1_cfg5_n1.json 2> gpurun_out/bench_r1_cfg5_n1.err; echo "cfg5 rc=$?"; cat gpurun_out/bench_r1_cfg5_n1.json
for c in cfg2 cfg1 cfg3 cfg3o cfg4; do
  timeout 600 python bench.py --workload $c --steps 20 --warmup 3 > gpurun_out/bench_r1_${c}_n1.json 2> gpurun_out/bench_r1_${c}_n1.err; echo "$c rc=$?"; cat gpurun_out/bench_r1_${c}_n1.json
done
timeout 600 python bench/blas1_bench.py > gpurun_out/r1_blas1.json 2>/dev/null; echo "blas1 rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_r1_reference_arm.json 2>/dev/null; cat gpurun_out/bench_r1_reference_arm.json
CMD5="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
CMD2="python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu --no-e2e"
CMD3="python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD5 > gpurun_out/plain5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hell_spmv_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1_cfg5_hell $CMD5 > gpurun_out/ncu5.log 2>&1; echo "ncu cfg5 rc=$?"
$CMD2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hdia_spmv_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1_cfg2_hdia $CMD2 > gpurun_out/ncu2.log 2>&1; echo "ncu cfg2 rc=$?"
$CMD3 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hell_spmv_kernel -s 3 -c 1 -f -o gpurun_out/prof_r1_cfg3_hell $CMD3 > gpurun_out/ncu3.log 2>&1; echo "ncu cfg3 rc=$?"
CMDL="python bench.py --steps 3 --warmup 3 --no-cpu --cg"
$CMDL > gpurun_out/plainL.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmv|reduce_kernel|ew_kernel|daxpby|halo' -c 400 --csv --log-file gpurun_out/launches_r1_cfg5.csv $CMDL > gpurun_out/ncuL.log 2>&1; echo "launch list rc=$?"
wc -l gpurun_out/launches_r1_cfg5.csv
