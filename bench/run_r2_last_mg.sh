#!/bin/bash
# the very last multi-GPU run: one cfg5 line with the final sources + the C-only driver
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-8}
O=gpurun_out/r2last$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?"; tail -1 $O/bench_cfg5.err
python - <<PY
import json
d = json.load(open("$O/bench_cfg5.json"))
print("ms/step", round(d["ms_per_step"], 5), "value", round(d["value"], 1), "frac", round(d["hbm_frac_of_peak"], 4), "verified", d.get("verified_vs_global_columns"), "status", d.get("device_status"), "clocks", d["clocks"],
      {k: (round(v["ms_per_iteration"], 4), round(v["frac_of_peak"], 4), v.get("kernels_per_iteration")) for k, v in (d.get("cg") or {}).items() if isinstance(v, dict)}, "e2e", (d.get("e2e") or {}).get("value"))
PY
gcc -O2 -fopenmp examples/mg_cg.c -Iinclude -I/usr/local/cuda/include -Lspgpu_b200/lib -lspgpu -Wl,-rpath,$PWD/spgpu_b200/lib -L/usr/local/cuda/lib64 -lcudart -lm -o /tmp/mg_cg && timeout 300 /tmp/mg_cg 512 $N 50 50 > $O/mg_cg_c_driver.txt 2>&1; echo "c driver rc=$?"; tail -4 $O/mg_cg_c_driver.txt
