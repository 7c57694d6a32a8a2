#!/bin/bash
# round-2 evidence run (1 GPU): full gpu test-suite, smoke, bench lines of all configurations (+ reference arm), BLAS-1 /
# stream-floor / conversion benchmarks, ncu --set full captures of the dominant kernels and the launch list.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2ev; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_r2_cfg5_n1.json 2> $O/bench_r2_cfg5_n1.err; echo "cfg5 rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_r2_reference_arm.json 2> $O/bench_r2_reference_arm.err; echo "ref rc=$?"
for c in cfg2 cfg2dia cfg1 cfg3 cfg3o cfg4; do
  timeout 600 python bench.py --workload $c --steps 20 --warmup 3 > $O/bench_r2_${c}_n1.json 2> $O/bench_r2_${c}_n1.err; echo "$c rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2ev/bench_r2_*.json")):
    try:
        d = json.load(open(f))
        rk = d.get("reference_kernels") or {}
        print(f.split("/")[-1], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 5), "frac", round((d.get("roofline") or {}).get("frac", 0), 4),
              "e2e", round((d.get("e2e") or {}).get("value", 0), 1), "cpu", round((d.get("cpu_baseline") or {}).get("value", 0), 2),
              "ref_kernels_x", round(rk.get("speedup_ours_vs_reference_kernels", 0), 2))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 300 python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu --no-e2e --no-ref-kernels --sweep "ellRows=1;ellRows=3;ellRows=4;ellRows=5;ellRows=2;ellRows=-1" 2>&1 | grep sweep | cut -c1-200
timeout 600 python bench/blas1_bench.py > $O/r2_blas1.json 2>/dev/null; echo "blas1 rc=$?"
timeout 300 python bench/stream_floor.py > $O/r2_stream_floor.json 2>/dev/null; echo "floor rc=$?"
timeout 600 python bench/conv_bench.py > $O/r2_conv.json 2>/dev/null; echo "conv rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-ref-kernels --no-cg"
cap() {  # name, kernel regex, bench args
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o $O/prof_r2_$1 $B $3 > $O/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
}
cap cfg5_hell hell_spmv_kernel ""
cap cfg2_hdia hdia_spmv_kernel "--workload cfg2"
cap cfg1_ell ell_spmv "--workload cfg1"
cap cfg3_hell hell_spmv_kernel "--workload cfg3"
cap cfg4_hell hell_spmv_kernel "--workload cfg4"
CMDL="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-ref-kernels"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmv|reduce_kernel|ew_kernel|axpby|cg_update|fold_partials|halo' -c 400 --csv --log-file $O/r2_cfg5_launches_ncu.csv $CMDL > $O/ncuL.log 2>&1; echo "launch list rc=$?"
wc -l $O/r2_cfg5_launches_ncu.csv; ls $O
