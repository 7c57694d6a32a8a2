#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spmv_gpu.py -m gpu -q -x -k "hdia" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu --no-e2e --sweep "hdiaVariant=5;hdiaVariant=0;hdiaVariant=5;hdiaVariant=0" > gpurun_out/b_cfg2.json 2> gpurun_out/b_cfg2.err
python -c "
import json;d=json.load(open('gpurun_out/b_cfg2.json'));print('cfg2 ms',d['ms_per_step'],'frac',d['roofline']['frac'])"; grep sweep gpurun_out/b_cfg2.err; grep -v "^frame" gpurun_out/b_cfg2.err | grep -v sweep | tail -3
