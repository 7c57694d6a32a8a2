#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for c in cfg3o cfg3; do
timeout 600 python bench.py --workload $c --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_$c.json 2> gpurun_out/b_$c.err; echo "$c rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/b_$c.json'));print('$c ms',d['ms_per_step'],'GF',d['value'],'frac',d['roofline']['frac'])"
done
timeout 600 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu --no-e2e --sweep "hdiaBlock=192;hdiaBlock=224;hdiaBlock=256;hdiaBlock=0" > gpurun_out/b_cfg2.json 2> gpurun_out/b_cfg2.err
python -c "
import json;d=json.load(open('gpurun_out/b_cfg2.json'));print('cfg2 ms',d['ms_per_step'],'frac',d['roofline']['frac'])"; grep sweep gpurun_out/b_cfg2.err
