#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload cfg1 --steps 20 --warmup 3 --no-cpu --no-e2e --sweep "hellVariant=3;hellVariant=1;hellVariant=2,hellBlock=64;hellVariant=2,hellBlock=192;hellVariant=2,hellBlock=256;hellVariant=0,hellBlock=0" > gpurun_out/b_cfg1.json 2> gpurun_out/b_cfg1.err
python -c "
import json;d=json.load(open('gpurun_out/b_cfg1.json'));print('cfg1 ms',d['ms_per_step'],'frac',d['roofline']['frac'])"; grep sweep gpurun_out/b_cfg1.err; grep -v "^frame" gpurun_out/b_cfg1.err | grep -v sweep | tail -3
