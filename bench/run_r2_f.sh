#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2f; mkdir -p $O
M=gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct,launch__occupancy_limit_registers,sm__maximum_warps_per_active_cycle_pct
timeout 900 ncu --metrics $M --clock-control none -k regex:'hell_spmv_kernel|spmv_halo_kernel' -c 16 --csv --log-file $O/halo_dot_probe.ncu.csv python bench/halo_dot_probe.py 512 1 > $O/halo_dot_probe.ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = collections.OrderedDict()
with open("gpurun_out/r2f/halo_dot_probe.ncu.csv") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.setdefault(r["ID"], {"k": r["Kernel Name"][:78]})[r["Metric Name"]] = r["Metric Value"]
for i, v in enumerate(rows.values()):
    if i % 4 != 3:
        continue
    print(v["k"])
    for k, x in v.items():
        if k != "k": print("     %-80s %s" % (k, x))
PY
