#!/bin/bash
# MatrixMarket harness: the gpu test, then one exported matrix through bench.py in every format
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mtx_harness_gpu.py -m gpu -q -x > gpurun_out/pytest_mtx.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_mtx.log
python - <<'PY'
from spgpu_b200 import generators as G, mmio
mmio.write_coo("/tmp/lap3d_64.mtx", G.laplace3d_7pt(64), "real", "symmetric")
PY
for f in hell ohell hdia ell dia; do
  timeout 600 python bench.py --matrix /tmp/lap3d_64.mtx --format $f --steps 10 --warmup 3 > gpurun_out/bench_r1_mtx_$f.json 2> gpurun_out/bench_r1_mtx_$f.err
  echo "mtx $f rc=$?"; cat gpurun_out/bench_r1_mtx_$f.json
done
