# Round 2: compute-sanitizer memcheck over the per-kernel conv tests (baton, t tiles, M-stacked weight gradients, two K chunks per
# stage, planar stores run in them) and over smoke(); racecheck over smoke().
set -x
mkdir -p gpurun_out/sanitize_r02
(timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -15) > gpurun_out/sanitize_r02/memcheck_kernels.log
(timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12) > gpurun_out/sanitize_r02/memcheck_smoke.log
(timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -25) > gpurun_out/sanitize_r02/racecheck_smoke.log
(timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "paper_config_train_step or tensor_core_path_agrees" 2>&1 | tail -12) > gpurun_out/sanitize_r02/memcheck_engine.log
tail -5 gpurun_out/sanitize_r02/*.log
true
