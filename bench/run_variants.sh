#!/bin/bash
# kernel-variant experiment run (1 GPU): the variant tests, then kernel-only sweeps on cfg2 / cfg1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_spmv_gpu.py -m gpu -q -x -k "hdia_kernel_variants or ell_short or ell_kernel_variants" > gpurun_out/pytest_variants.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_variants.log
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu --no-e2e \
  --sweep "hdiaVariant=0,hdiaBlock=0;hdiaVariant=0,hdiaBlock=9;hdiaVariant=0,hdiaBlock=64;hdiaVariant=0,hdiaBlock=176;hdiaVariant=0,hdiaBlock=192;hdiaVariant=0,hdiaBlock=0" \
  > gpurun_out/sweep_cfg2.json 2> gpurun_out/sweep_cfg2.err; echo "cfg2 rc=$?"; grep sweep gpurun_out/sweep_cfg2.err
timeout 300 python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu --no-e2e \
  --sweep "ellRows=0;ellRows=1;ellRows=2;ellRows=0" \
  > gpurun_out/sweep_cfg1.json 2> gpurun_out/sweep_cfg1.err; echo "cfg1 rc=$?"; grep sweep gpurun_out/sweep_cfg1.err
