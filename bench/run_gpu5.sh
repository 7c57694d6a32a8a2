#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --sweep "hellVariant=3;hellVariant=2;hellVariant=1;hellVariant=0" > gpurun_out/b5.json 2> gpurun_out/b5.err; echo "cfg5 rc=$?"
python -c "import json;d=json.load(open('gpurun_out/b5.json'));print('cfg5',d['ms_per_step'],d['roofline']['frac'])"; grep sweep gpurun_out/b5.err; grep -v "^frame" gpurun_out/b5.err | grep -v sweep | tail -3
for c in cfg2 cfg1 cfg3; do
timeout 300 python bench.py --workload $c --steps 20 --warmup 3 --no-cpu --no-e2e --sweep "hellVariant=1;hellVariant=3;hellVariant=0,hellBlock=64;hellVariant=0,hellBlock=0" > gpurun_out/b_$c.json 2> gpurun_out/b_$c.err; echo "$c rc=$?"
python -c "import json;d=json.load(open('gpurun_out/b_$c.json'));print('$c',d['ms_per_step'],d['roofline']['frac'])"; grep sweep gpurun_out/b_$c.err; grep -v "^frame" gpurun_out/b_$c.err | grep -v sweep | tail -3
done
