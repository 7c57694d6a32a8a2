#!/bin/bash
# round 2, third 1-GPU pass: full test-suite, cfg3 / cfg3o / cfg4 with the 64-byte-granularity loads (+ ncu DRAM bytes),
# launch list of the default bench, reductions after the new defaults
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $O/pytest_gpu.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
for c in cfg3 cfg3o cfg4; do
  timeout 400 python bench.py --workload $c --steps 20 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c rc=$?"
  python -c "
import json; d=json.load(open('$O/bench_$c.json')); print('$c', 'ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'gflops', d['value'], 'ref_kernels', (d.get('reference_kernels') or {}).get('kernel_ms'))"
  timeout 300 ncu --metrics $M --clock-control none -k regex:hell_spmv_kernel -s 3 -c 1 --csv --log-file $O/$c.ncu.csv python bench.py --workload $c --steps 2 --warmup 3 --no-cpu --no-e2e --no-ref-kernels > $O/$c.ncu.log 2>&1; echo "ncu rc=$?"
  grep -E "dram__bytes_read|gpu__time|lts__t_sectors|l1tex__t" $O/$c.ncu.csv | awk -F'","' '{print $5, $(NF-2), $NF}'
done
CMDL="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-ref-kernels"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmv|reduce_kernel|ew_kernel|axpby|cg_update|fold_partials|halo|scal' -c 400 --csv --log-file $O/launches_cfg5.csv $CMDL > $O/ncuL.log 2>&1; echo "launch list rc=$?"; wc -l $O/launches_cfg5.csv
python - <<'PY'
import csv, collections
agg = collections.OrderedDict()
with open("gpurun_out/r2c/launches_cfg5.csv") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    k = r["Kernel Name"][:90]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += float(r["Metric Value"].replace(",", "")) / 1e3
for k, (n, us) in agg.items():
    print(f"{n:4d} x {us / n:10.1f} us  {k}")
PY
timeout 600 python bench/blas1_bench.py > $O/blas1.json 2> $O/blas1.err; echo "blas1 rc=$?"; python -c "
import json; d=json.load(open('$O/blas1.json'))
for k,v in d.items():
    if isinstance(v,dict) and 'frac' in v: print(f'{k:45s} {v[\"frac\"]:.3f}')"
