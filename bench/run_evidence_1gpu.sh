#!/bin/bash
# round-1 evidence run (1 GPU): full gpu test-suite, smoke, bench lines of all configurations, BLAS-1 /
# stream-floor / conversion benchmarks, the reference arm, ncu captures and the launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --cg > gpurun_out/bench_r1_cfg5_n1.json 2> gpurun_out/bench_r1_cfg5_n1.err; echo "cfg5 rc=$?"; cat gpurun_out/bench_r1_cfg5_n1.json
for c in cfg2 cfg2dia cfg1 cfg3 cfg3o cfg4; do
  timeout 600 python bench.py --workload $c --steps 20 --warmup 3 > gpurun_out/bench_r1_${c}_n1.json 2> gpurun_out/bench_r1_${c}_n1.err; echo "$c rc=$?"; cat gpurun_out/bench_r1_${c}_n1.json
done
timeout 600 python bench/blas1_bench.py > gpurun_out/r1_blas1.json 2>/dev/null; echo "blas1 rc=$?"
timeout 300 python bench/stream_floor.py > gpurun_out/r1_stream_floor.json 2>/dev/null; echo "floor rc=$?"
timeout 600 python bench/conv_bench.py > gpurun_out/r1_conv.json 2>/dev/null; echo "conv rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_r1_reference_arm.json 2>/dev/null; cat gpurun_out/bench_r1_reference_arm.json
timeout 600 python -m pytest tests/test_reference_perf_gpu.py -m gpu -q -s > gpurun_out/ref_perf.log 2>&1; echo "ref perf rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
cap() {  # name, kernel regex, bench args
  $B $3 > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/prof_r1_$1 $B $3 > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
}
cap cfg5_hell hell_spmv_kernel ""
cap cfg2_hdia hdia_spmv_kernel "--workload cfg2"
cap cfg1_ell ell_spmv "--workload cfg1"
cap cfg3_hell hell_spmv_kernel "--workload cfg3"
cap cfg3o_hell hell_spmv_kernel "--workload cfg3o"
CMDL="python bench.py --steps 3 --warmup 3 --no-cpu --cg"
$CMDL > gpurun_out/plainL.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmv|reduce_kernel|ew_kernel|daxpby|halo|dcg' -c 400 --csv --log-file gpurun_out/launches_r1_cfg5.csv $CMDL > gpurun_out/ncuL.log 2>&1; echo "launch list rc=$?"
wc -l gpurun_out/launches_r1_cfg5.csv
