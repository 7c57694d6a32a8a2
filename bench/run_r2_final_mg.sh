#!/bin/bash
# closing multi-GPU run: cfg5 under torchrun with the final defaults (programmatic dependent launch on), the same with it off
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-8}
O=gpurun_out/r2fin$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?"; tail -2 $O/bench_cfg5.err
SPGPU_PDL=0 timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-e2e > $O/bench_cfg5_pdl0.json 2> $O/bench_cfg5_pdl0.err; echo "cfg5 pdl=0 rc=$?"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-e2e > $O/bench_cfg5_b.json 2> $O/bench_cfg5_b.err; echo "cfg5 (again) rc=$?"
python - <<PY
import json
for f in ("bench_cfg5", "bench_cfg5_pdl0", "bench_cfg5_b"):
    try:
        d = json.load(open("$O/" + f + ".json"))
        print(f, "ms/step", round(d["ms_per_step"], 5), "value", round(d["value"], 1), "b2b", d["roofline"].get("kernel_back_to_back_ms"), "frac", round(d["hbm_frac_of_peak"], 4),
              "verified", d.get("verified_vs_global_columns"), "status", d.get("device_status"), "traffic", d["roofline"].get("traffic"),
              {k: (round(v["ms_per_iteration"], 4), round(v["frac_of_peak"], 4)) for k, v in (d.get("cg") or {}).items() if isinstance(v, dict)},
              "e2e", (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "no json", e)
PY
if [ "$2" == "cdriver" ]; then
gcc -O2 -fopenmp examples/mg_cg.c -Iinclude -I/usr/local/cuda/include -Lspgpu_b200/lib -lspgpu -Wl,-rpath,$PWD/spgpu_b200/lib -L/usr/local/cuda/lib64 -lcudart -lm -o /tmp/mg_cg && timeout 600 /tmp/mg_cg 512 $N 50 50 > $O/mg_cg_c_driver.txt 2>&1; echo "c driver rc=$?"; cat $O/mg_cg_c_driver.txt
fi
