set -x
mkdir -p gpurun_out/s17
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/s17/pytest.log
timeout 200 python tools/step_profile.py > gpurun_out/s17/step_profile.txt 2>&1
timeout 300 python tools/conv_shapes.py --time "32 32 32 192 192 3 2" "32 64 32 192 192 3 2" "32 64 64 96 96 3 2" "32 128 128 48 48 3 2" "32 512 512 12 12 3 2" "32 256 256 24 24 3 2" "32 1024 1024 6 6 3 2" "32 64 32 192 192 1 2" > gpurun_out/s17/wgrad_shapes.txt 2>&1
(timeout 300 python bench.py 2>gpurun_out/s17/bench.err | tail -2) > gpurun_out/s17/bench.log
du -sh gpurun_out
