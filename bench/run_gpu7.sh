#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --cg --sweep "hellBlock=192;hellBlock=256;hellBlock=64;hellBlock=0" > gpurun_out/b5.json 2> gpurun_out/b5.err; echo "cfg5 rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/b5.json'));print('cfg5 ms',d['ms_per_step'],'frac',d['roofline']['frac']);print('e2e',json.dumps(d['e2e']));print('cg',json.dumps(d['cg']))"
grep sweep gpurun_out/b5.err; grep -v "^frame" gpurun_out/b5.err | grep -v sweep | tail -3
timeout 600 python bench.py --workload cfg4 --steps 20 --warmup 3 --no-cpu > gpurun_out/b4.json 2> gpurun_out/b4.err; echo "cfg4 rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/b4.json'));print('cfg4 ms',d['ms_per_step'],'frac',d['roofline']['frac']);print('e2e',json.dumps(d['e2e']))"
