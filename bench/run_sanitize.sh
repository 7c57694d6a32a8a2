#!/bin/bash
# compute-sanitizer memcheck over the SpMV entry points (every format / type / kernel variant) on small
# matrices; torch's caching allocator is disabled so every tensor is its own cudaMalloc and an
# out-of-bounds access of a kernel cannot hide inside a cached segment.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
SEL="random_ragged or variants or empty_matrix or hack_sizes or ell_without"
timeout 1200 python -m pytest tests/test_spmv_gpu.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitize_plain.log 2>&1 && \
timeout 3000 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/r1_memcheck.log \
    python -m pytest tests/test_spmv_gpu.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitize_pytest.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/sanitize_plain.log; tail -5 gpurun_out/sanitize_pytest.log; tail -8 gpurun_out/r1_memcheck.log
