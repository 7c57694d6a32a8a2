#!/bin/bash
# multi-GPU scaling evidence: N ranks (torchrun), cfg5 512^3, halo modes, CG
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}; MODES=${2:-"fused push nccl"}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for mode in $MODES; do
  echo "== N=$N full 512^3 halo=$mode"
  timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --halo $mode --verify > gpurun_out/mg${N}_$mode.json 2> gpurun_out/mg${N}_$mode.err; echo "rc=$?"
  grep -E "verify|rror|timed out" gpurun_out/mg${N}_$mode.err | head -5
  python -c "import json;d=json.load(open('gpurun_out/mg${N}_$mode.json'));print(d['n_gpus'],d['config']['parallelism'],'ms',d['ms_per_step'],'GF',d['value'],'frac',d['hbm_frac_of_peak'],'kernel_ms',d['roofline']['kernel_ms'],'e2e',d['e2e'] and d['e2e']['ms_per_step'],'launches',d['gpu_launches'])"
done
echo "== N=$N CG (fused halo)"
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --halo fused --cg --no-e2e > gpurun_out/mg${N}_cg.json 2> gpurun_out/mg${N}_cg.err; echo "rc=$?"
grep -E "rror|timed out" gpurun_out/mg${N}_cg.err | head -5
python -c "import json;d=json.load(open('gpurun_out/mg${N}_cg.json'));print(json.dumps(d['cg']))"
