#!/bin/bash
# round 2, two GPUs: the multi-rank tests, then cfg5 / cfg2 under torchrun with the per-exchange trace
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
O=gpurun_out/r2n$N; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
if [ "$2" != "notests" ]; then
timeout 1200 python -m pytest tests/test_mg_multi_gpu.py tests/test_mg_capi_gpu.py tests/test_c_driver_gpu.py tests/test_mg_gpu.py tests/test_krylov_gpu.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -30 $O/pytest.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
SPGPU_BENCH_TRACE=1 timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-e2e > $O/bench_cfg5_trace.json 2> $O/bench_cfg5_trace.err; echo "cfg5 trace rc=$?"
grep -E "CG phases|CG exchange" $O/bench_cfg5_trace.err; grep -E "^rank [0-9]+ seq" $O/bench_cfg5_trace.err | tail -4
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?"; tail -3 $O/bench_cfg5.err
python - <<PY
import json
for f in ("bench_cfg5_trace", "bench_cfg5"):
    try:
        d = json.load(open("$O/" + f + ".json"))
        print(f, "ms/step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "overhead_us", (d["ms_per_step"] - d["roofline"]["kernel_ms"]) * 1e3,
              "frac", d["hbm_frac_of_peak"], "verified", d.get("verified_vs_global_columns"), "status", d.get("device_status"))
        if "cg" in d:
            print({k: (v["ms_per_iteration"], round(v["frac_of_peak"], 4), v.get("kernels_per_iteration")) for k, v in d["cg"].items() if isinstance(v, dict)})
        if d.get("e2e"): print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    except Exception as e:
        print(f, "no json", e)
PY
timeout 600 $TR bench.py --gpus $N --workload cfg2 --steps 20 --warmup 5 --no-e2e > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"; cat $O/bench_cfg2.json | python -c "
import json,sys; d=json.load(sys.stdin); print('cfg2 ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'verified', d.get('verified_vs_global_columns'))"
