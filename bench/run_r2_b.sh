#!/bin/bash
# round 2, second 1-GPU pass: tests, sector-granularity probe (plain + ncu), cfg5 bench + launch list, reduction sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2b; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 $O/pytest_gpu.log
timeout 120 bench/probes/sector_probe > $O/sector_probe.jsonl 2>&1; echo "probe rc=$?"; cat $O/sector_probe.jsonl
timeout 600 ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum --clock-control none --csv --log-file $O/sector_probe.ncu.csv bench/probes/sector_probe > $O/sector_probe.ncu.log 2>&1; echo "probe ncu rc=$?"
python - <<'PY'
import csv, collections
rows = collections.OrderedDict()
try:
    with open("gpurun_out/r2b/sector_probe.ncu.csv") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.setdefault(r["ID"], {"k": r["Kernel Name"]})[r["Metric Name"]] = r["Metric Value"]
    seen = set()
    for i, (k, v) in enumerate(rows.items()):
        if i % 2 == 0:      # first launch of each pair is the warm-up
            continue
        print(v["k"][:40], "dramMB", int(v.get("dram__bytes_read.sum","0").replace(",",""))//1000000, "ltsRdSect(M)", int(v.get("lts__t_sectors_srcunit_tex_op_read.sum","0").replace(",",""))//1000000, "l1Sect(M)", int(v.get("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum","0").replace(",",""))//1000000, "us", int(v.get("gpu__time_duration.sum","0").replace(",",""))//1000)
except Exception as e:
    print("parse failed", e)
PY
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?"; python -c "
import json; d=json.load(open('$O/bench_cfg5.json')); print('value', d['value'], 'frac', d['roofline']['frac'], 'status', d['device_status']); print(json.dumps(d['cg'], indent=0)); print(d['cpu_baseline'])"; tail -3 $O/bench_cfg5.err
CMDL="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-ref-kernels"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg5.csv $CMDL > $O/ncuL.log 2>&1; echo "launch list rc=$?"; wc -l $O/launches_cfg5.csv
python - <<'PY'
import csv, collections
agg = collections.OrderedDict()
with open("gpurun_out/r2b/launches_cfg5.csv") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    k = r["Kernel Name"][:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += float(r["Metric Value"].replace(",", "")) / 1e3
for k, (n, us) in agg.items():
    print(f"{n:4d} x {us / n:10.1f} us  {k}")
PY
timeout 900 python bench/blas1_bench.py --sweep > $O/blas1.json 2> $O/blas1.err; echo "blas1 rc=$?"; python -c "
import json; d=json.load(open('$O/blas1.json'))
for k,v in d.items():
    if isinstance(v,dict) and 'frac' in v: print(f'{k:45s} {v[\"frac\"]:.3f}')
for k,v in d.get('reduction_sweep_frac_of_peak',{}).items(): print(k, v)"
