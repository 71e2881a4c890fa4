# compute-sanitizer memcheck over the per-kernel conv tests (tcgen05 + CUDA-core paths) and racecheck over smoke().
set -x
mkdir -p gpurun_out/sanitize
(timeout 75 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -15) > gpurun_out/sanitize/memcheck_kernels.log
(timeout 55 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -25) > gpurun_out/sanitize/racecheck_smoke.log
true
