#!/bin/bash
# round-2 closing run (1 GPU): the C multi-GPU API tests first (new HDIA paths), then the whole gpu suite, smoke, and the
# bench lines the driver will take (own arm + reference arm) plus the cfg3 / cfg2 lines with the final bench.py.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2fin; mkdir -p $O
timeout 600 python -m pytest tests/test_mg_capi_gpu.py -q -x > $O/pytest_mg_capi.log 2>&1; echo "mg capi rc=$?"; tail -15 $O/pytest_mg_capi.log
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
timeout 600 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "ref rc=$?"
for c in cfg3 cfg2; do
  timeout 600 python bench.py --workload $c --steps 20 --warmup 3 > $O/bench_${c}.json 2> $O/bench_${c}.err; echo "$c rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2fin/bench_*.json")):
    try:
        d = json.load(open(f))
        rk = d.get("reference_kernels") or {}
        print(f.split("/")[-1], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 5), "frac", round((d.get("roofline") or {}).get("frac", 0), 4),
              "e2e", round((d.get("e2e") or {}).get("value", 0), 1), "cpu", round((d.get("cpu_baseline") or {}).get("value", 0), 2),
              "ref_kernels_x", round(rk.get("speedup_ours_vs_reference_kernels", 0), 2), "status", d.get("device_status"))
    except Exception as e:
        print(f, "unreadable", e)
PY
