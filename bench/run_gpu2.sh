#!/bin/bash
# second GPU call: full gpu test-suite + first bench lines of every configuration
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --sweep "hellBlock=64;hellBlock=256;hellBlock=512;hellBlock=128" > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "bench cfg5 rc=$?"
cat gpurun_out/bench_cfg5.json; grep sweep gpurun_out/bench_cfg5.err; tail -3 gpurun_out/bench_cfg5.err
for c in cfg2 cfg1 cfg3 cfg4; do
  python bench.py --workload $c --steps 20 --warmup 3 --no-cpu --sweep "hellBlock=64,hdiaBlock=64;hellBlock=256,hdiaBlock=256;hellBlock=128,hdiaBlock=128" > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c rc=$?"
  cat gpurun_out/bench_$c.json; grep sweep gpurun_out/bench_$c.err; tail -2 gpurun_out/bench_$c.err
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
