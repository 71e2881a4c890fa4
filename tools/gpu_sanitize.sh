# compute-sanitizer over the sample-prep / post-processing kernels (memcheck + racecheck) and over smoke() (memcheck).
set -x
mkdir -p gpurun_out/sanitize
K='not full_size and not seg_dataset_ensemble'
(timeout 60 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_prepost_gpu.py -m gpu -q -x -k "$K" 2>&1 | tail -12) > gpurun_out/sanitize/memcheck_prepost.log
(timeout 60 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_prepost_gpu.py -m gpu -q -x -k "$K" 2>&1 | tail -12) > gpurun_out/sanitize/racecheck_prepost.log
(timeout 50 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12) > gpurun_out/sanitize/memcheck_smoke.log
tail -4 gpurun_out/sanitize/*.log
