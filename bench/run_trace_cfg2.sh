cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
for mode in fused push; do
  SPGPU_BENCH_TRACE=1 timeout 300 $TR bench.py --gpus $N --workload cfg2 --steps 20 --warmup 3 --halo $mode --no-cpu --no-e2e > gpurun_out/trace_$mode.json 2> gpurun_out/trace_$mode.err
  echo "$mode rc=$?"; grep "step ms" gpurun_out/trace_$mode.err; python -c "import json; d=json.load(open('gpurun_out/trace_$mode.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'])"
done
