#!/bin/bash
# programmatic dependent launch on a partition: multi-rank tests with it on, then cfg5 / cfg2 under torchrun, off and on
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
O=gpurun_out/r2pdl$N; mkdir -p $O
if [ "$2" != "notests" ]; then
SPGPU_PDL=1 timeout 900 python -m pytest tests/test_mg_multi_gpu.py tests/test_mg_capi_gpu.py tests/test_c_driver_gpu.py tests/test_mg_gpu.py tests/test_krylov_gpu.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest pdl=1 rc=$?"; tail -5 $O/pytest.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
for v in 0 1 0 1; do
  SPGPU_PDL=$v timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu > $O/bench_cfg5_pdl${v}.json 2> $O/bench_cfg5_pdl${v}.err; echo "cfg5 pdl=$v rc=$?"
  python - <<PY
import json
d = json.load(open("$O/bench_cfg5_pdl${v}.json"))
print("cfg5 pdl=$v ms/step", round(d["ms_per_step"], 5), "b2b", d["roofline"].get("kernel_back_to_back_ms"), "frac", round(d["hbm_frac_of_peak"], 4), "verified", d.get("verified_vs_global_columns"), "status", d.get("device_status"),
      {k: (round(v["ms_per_iteration"], 4), round(v["frac_of_peak"], 4)) for k, v in (d.get("cg") or {}).items() if isinstance(v, dict)})
PY
done
for v in 0 1; do
  SPGPU_PDL=$v timeout 600 $TR bench.py --gpus $N --workload cfg2 --steps 20 --warmup 5 --no-e2e --no-cpu > $O/bench_cfg2_pdl${v}.json 2> $O/bench_cfg2_pdl${v}.err; echo "cfg2 pdl=$v rc=$?"
  python -c "
import json; d=json.load(open('$O/bench_cfg2_pdl${v}.json')); print('cfg2 pdl=$v ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'verified', d.get('verified_vs_global_columns'), 'status', d.get('device_status'))"
done
