#!/bin/bash
# per-step, per-rank times of the partitioned cfg2 SpMV (diagnosing rank skew): N ranks, given halo modes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-4}; MODES=${2:-"fused push"}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
for mode in $MODES; do
  SPGPU_BENCH_TRACE=1 timeout 300 $TR bench.py --gpus $N --workload cfg2 --steps 20 --warmup 3 --halo $mode --no-cpu --no-e2e > gpurun_out/trace_n${N}_$mode.json 2> gpurun_out/trace_n${N}_$mode.err
  echo "$mode rc=$?"; grep -o "rank [0-9] step ms:[ 0-9.]*" gpurun_out/trace_n${N}_$mode.err; python -c "import json; d=json.load(open('gpurun_out/trace_n${N}_$mode.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'])"
done
