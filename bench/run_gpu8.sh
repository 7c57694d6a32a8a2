#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
cat gpurun_out/ref_vs_ours.json
for c in cfg3o cfg3; do
timeout 600 python bench.py --workload $c --steps 20 --warmup 3 --no-cpu > gpurun_out/b_$c.json 2> gpurun_out/b_$c.err; echo "$c rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/b_$c.json'));print('$c ms',d['ms_per_step'],'GF',d['value'],'frac',d['roofline']['frac'])"
grep -v "^frame" gpurun_out/b_$c.err | tail -3
done
