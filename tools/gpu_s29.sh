mkdir -p gpurun_out/s29
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/s29/pytest.log
timeout 200 python tools/step_profile.py > gpurun_out/s29/step_profile.txt 2>&1
(timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1) > gpurun_out/s29/bench.log
