# Round 2, late: compute-sanitizer over the kernels added after the first sanitizer pass -- tensor-core heads forward / backward
# (+ their loss-fused variants), vectorised loss gradient, read-and-clear unpack + gap zeroing, kw-stacked halo MMAs.
set -x
mkdir -p gpurun_out/sanitize_r02b
(timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_loss_gpu.py tests/test_engine_gpu.py -m gpu -q -x -k "heads or loss or repeated_backward or golden_case" 2>&1 | tail -12) > gpurun_out/sanitize_r02b/memcheck_heads_loss.log
(timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "halo" 2>&1 | tail -12) > gpurun_out/sanitize_r02b/memcheck_halo_stack.log
(timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_loss_gpu.py -m gpu -q -x -k "inside_the_heads" 2>&1 | tail -15) > gpurun_out/sanitize_r02b/racecheck_heads_loss.log
(timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12) > gpurun_out/sanitize_r02b/memcheck_smoke.log
tail -4 gpurun_out/sanitize_r02b/*.log
true
