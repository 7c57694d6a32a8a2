cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests/test_spmv_gpu.py tests/test_large_gpu.py -m gpu -q -x 2>&1 | tail -4
for c in cfg3o cfg3; do
python bench.py --workload $c --steps 20 --warmup 3 --no-cpu --no-e2e --sweep "hellSplit=-1;hellSplit=1;hellSplit=0" 2> gpurun_out/b_$c.err | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$c', d['ms_per_step'], d['value'], d['roofline']['frac'], d['gpu_launches'])"; grep sweep gpurun_out/b_$c.err; grep -v "^frame" gpurun_out/b_$c.err | grep -v sweep | tail -2
done
