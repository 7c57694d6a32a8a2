#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2d; mkdir -p $O
timeout 300 python bench/halo_dot_probe.py 512 10 > $O/halo_dot_probe.json 2> $O/halo_dot_probe.err; echo "probe rc=$?"; cat $O/halo_dot_probe.json; tail -3 $O/halo_dot_probe.err
M=gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,smsp__warp_issue_stalled_barrier_per_warp_active.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_membar_per_warp_active.pct
timeout 600 ncu --metrics $M --clock-control none -k regex:'hell_spmv_kernel|spmv_halo_kernel' -s 12 -c 40 --csv --log-file $O/halo_dot_probe.ncu.csv python bench/halo_dot_probe.py 512 1 > $O/halo_dot_probe.ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = collections.OrderedDict()
with open("gpurun_out/r2d/halo_dot_probe.ncu.csv") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.setdefault(r["ID"], {"k": r["Kernel Name"][:75]})[r["Metric Name"]] = r["Metric Value"]
for v in rows.values():
    print(v["k"])
    print("   ", {k.split("__")[-1][:40]: x for k, x in v.items() if k != "k"})
PY
# cfg3: full ncu capture of the HELL kernel (what bounds it after the 64-byte granularity change?)
B="python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-e2e --no-ref-kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hell_spmv_kernel -s 3 -c 1 -f -o $O/prof_cfg3 $B > $O/ncu_cfg3.log 2>&1; echo "ncu cfg3 rc=$?"
ls -la $O
