#!/bin/bash
# cfg2 (HDIA) row-sharded over N GPUs: fused, push and nccl halo modes, verified against the full matrix
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
for mode in fused push nccl; do
  timeout 600 $TR bench.py --gpus $N --workload cfg2 --steps 20 --warmup 3 --halo $mode --verify \
    > gpurun_out/bench_r1_cfg2_n${N}_${mode}.json 2> gpurun_out/bench_r1_cfg2_n${N}_${mode}.err
  echo "cfg2 N=$N $mode rc=$?"; grep -h "verify" gpurun_out/bench_r1_cfg2_n${N}_${mode}.err; cat gpurun_out/bench_r1_cfg2_n${N}_${mode}.json
done
# regression check of the refactored fused HELL kernel on cfg5
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --halo fused --verify --no-cpu \
    > gpurun_out/bench_r1_cfg5_n${N}_fused.json 2> gpurun_out/bench_r1_cfg5_n${N}_fused.err
echo "cfg5 N=$N fused rc=$?"; grep -h "verify" gpurun_out/bench_r1_cfg5_n${N}_fused.err; cat gpurun_out/bench_r1_cfg5_n${N}_fused.json
