#!/bin/bash
# programmatic dependent launch A/B (tuning key pdl / env SPGPU_PDL): the whole gpu suite with it on, then timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2pdl; mkdir -p $O
SPGPU_PDL=1 timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_pdl1.log 2>&1; echo "pytest pdl=1 rc=$?"; tail -5 $O/pytest_pdl1.log
timeout 600 python -m pytest tests/test_krylov_gpu.py tests/test_blas1_gpu.py tests/test_mg_capi_gpu.py -q -x > $O/pytest_pdl0.log 2>&1; echo "pytest pdl=0 rc=$?"; tail -3 $O/pytest_pdl0.log
for v in 0 1 0 1; do
  for p in 8 2; do
    echo "== probe parts=$p pdl=$v"; SPGPU_PDL=$v PROBE_PARTS=$p timeout 300 python bench/halo_dot_probe.py 512 20 2>&1 | tr -d '\n'; echo
  done
done
for v in 0 1; do
  SPGPU_PDL=$v timeout 600 python bench.py --size 256 --steps 40 --warmup 5 --no-cpu --no-e2e --no-ref-kernels > $O/bench_256_pdl$v.json 2> $O/bench_256_pdl$v.err; echo "256^3 pdl=$v rc=$?"
  SPGPU_PDL=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-ref-kernels > $O/bench_512_pdl$v.json 2> $O/bench_512_pdl$v.err; echo "512^3 pdl=$v rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2pdl/bench_*.json")):
    try:
        d = json.load(open(f)); cg = d.get("cg") or {}
        print(f.split("/")[-1], "ms", round(d["ms_per_step"], 5), "b2b", d["roofline"].get("kernel_back_to_back_ms"), "kernel", round(d["roofline"]["kernel_ms"], 5),
              "cg blocking/device/graph", [round((cg.get(k) or {}).get("ms_per_iteration", 0), 4) for k in ("blocking", "device", "graph")],
              "rr", [(cg.get(k) or {}).get("residual_norm2_after") for k in ("blocking", "device", "graph")])
    except Exception as e:
        print(f, "unreadable", e)
PY
