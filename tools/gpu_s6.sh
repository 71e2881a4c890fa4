set -x
mkdir -p gpurun_out/s6
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/s6/pytest.log
timeout 300 python tools/step_profile.py > gpurun_out/s6/step_profile.txt 2>&1
timeout 300 python tools/profile_layers.py 32 192 bf16 > gpurun_out/s6/layers.txt 2>&1
(timeout 300 python bench.py 2>&1 | tail -2) > gpurun_out/s6/bench.log
du -sh gpurun_out
