#!/bin/bash
# DRAM traffic of the fused SpMV + halo kernel at the slab sizes of 2 / 4 / 8 ranks (cfg5, one GPU, emulated neighbours whose
# flags say "arrived": ncu must not be run on a multi-rank command), ncu --set full; plus the +dot variant at the 8-rank slab.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2tr; mkdir -p $O
timeout 300 python -m pytest tests/test_mg_capi_gpu.py -q -x -k "fewer_hacks or hdia" > $O/pytest_edge.log 2>&1; echo "edge rc=$?"; tail -5 $O/pytest_edge.log
for p in 2 4 8; do
  PROBE_PARTS=$p timeout 300 python bench/halo_dot_probe.py 512 10 > $O/probe_parts$p.json 2>&1; echo "probe $p rc=$?"; cat $O/probe_parts$p.json
  # launches of spmv_halo_kernel in the probe: 13 x (halo), 13 x (dot, no halo), 13 x (halo + dot)
  PROBE_PARTS=$p ncu --set full --clock-control none --import-source on -k regex:spmv_halo_kernel -s 5 -c 1 -f -o $O/prof_r2_cfg5_halo_n$p \
     python bench/halo_dot_probe.py 512 10 > $O/ncu_halo_n$p.log 2>&1; echo "ncu halo n$p rc=$?"
done
PROBE_PARTS=8 ncu --set full --clock-control none --import-source on -k regex:spmv_halo_kernel -s 31 -c 1 -f -o $O/prof_r2_cfg5_halodot_n8 \
   python bench/halo_dot_probe.py 512 10 > $O/ncu_halodot_n8.log 2>&1; echo "ncu halodot n8 rc=$?"
ls -la $O
