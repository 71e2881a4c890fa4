set -x
mkdir -p gpurun_out/s22
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/s22/pytest.log
timeout 200 python tools/step_profile.py > gpurun_out/s22/step_profile.txt 2>&1
(timeout 300 python bench.py 2>gpurun_out/s22/bench.err | tail -2) > gpurun_out/s22/bench.log
du -sh gpurun_out
