#!/bin/bash
# multi-GPU: N ranks (torchrun), cfg5, halo modes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== small verify (n=128) push"; timeout 300 $TR bench.py --gpus $N --size 128 --steps 5 --warmup 3 --no-e2e --verify --halo fused > gpurun_out/mg${N}_small_push.json 2> gpurun_out/mg${N}_small_push.err; echo "rc=$?"; grep -E "verify|Error|error|timed out" gpurun_out/mg${N}_small_push.err | head -5; cat gpurun_out/mg${N}_small_push.json | cut -c1-400
echo "== small verify (n=128) nccl"; timeout 300 $TR bench.py --gpus $N --size 128 --steps 5 --warmup 3 --no-e2e --verify --halo nccl > gpurun_out/mg${N}_small_nccl.json 2> gpurun_out/mg${N}_small_nccl.err; echo "rc=$?"; grep -E "verify|Error|error" gpurun_out/mg${N}_small_nccl.err | head -5; cat gpurun_out/mg${N}_small_nccl.json | cut -c1-400
for mode in "fused" "push" "nccl"; do
  tag=$(echo $mode | tr -d ' -')
  echo "== full 512^3 halo=$mode"
  timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --halo $mode --verify > gpurun_out/mg${N}_$tag.json 2> gpurun_out/mg${N}_$tag.err; echo "rc=$?"
  grep -E "verify|rror|timed out" gpurun_out/mg${N}_$tag.err | head -5
  python -c "import json;d=json.load(open('gpurun_out/mg${N}_$tag.json'));print(d['n_gpus'],d['config']['parallelism'],'ms',d['ms_per_step'],'GF',d['value'],'frac',d['hbm_frac_of_peak'],'e2e',d['e2e'] and d['e2e']['ms_per_step'],'launches',d['gpu_launches'])"
done
echo "== CG (fused halo)"
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --halo fused --cg --no-e2e > gpurun_out/mg${N}_cg.json 2> gpurun_out/mg${N}_cg.err; echo "rc=$?"
grep -E "rror|timed out" gpurun_out/mg${N}_cg.err | head -5
python -c "import json;d=json.load(open('gpurun_out/mg${N}_cg.json'));print(json.dumps(d['cg']))"
