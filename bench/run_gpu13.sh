#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_large_gpu.py tests/test_conv_device_gpu.py -m gpu -q -x > gpurun_out/pytest_large.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_large.log
timeout 600 python bench.py --workload cfg2dia --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_r1_cfg2dia_n1.json 2> gpurun_out/b_cfg2dia.err; echo "cfg2dia rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_r1_cfg2dia_n1.json'));print('cfg2dia ms',d['ms_per_step'],'GF',d['value'],'frac',d['roofline']['frac'])"; grep -v "^frame" gpurun_out/b_cfg2dia.err | tail -3
