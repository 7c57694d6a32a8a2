#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2e; mkdir -p $O
timeout 300 python bench/halo_dot_probe.py 512 10 > $O/halo_dot_probe.json 2> $O/halo_dot_probe.err; echo "probe rc=$?"; cat $O/halo_dot_probe.json; tail -3 $O/halo_dot_probe.err
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu --no-e2e --no-ref-kernels --no-verify \
  --sweep "hellSplit=1;hellSplit=-1,hellLongFactor=2;hellLongFactor=8;hellLongFactor=16;hellLongFactor=4,hellBlock=192;hellBlock=64;hellBlock=0,hellSplit=1,hellLongFactor=2;hellLongFactor=8;hellSplit=0,hellLongFactor=4,hellVariant=1;hellVariant=3" \
  > $O/cfg3_sweep.json 2> $O/cfg3_sweep.err; echo "sweep rc=$?"; grep sweep $O/cfg3_sweep.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-ref-kernels --no-cpu > $O/bench_cfg5.json 2> $O/bench_cfg5.err; python -c "
import json; d=json.load(open('$O/bench_cfg5.json')); print('cfg5', d['value'], d['roofline']['frac']); print({k: (v['ms_per_iteration'], round(v['frac_of_peak'],4)) for k,v in d['cg'].items() if isinstance(v, dict)})"
