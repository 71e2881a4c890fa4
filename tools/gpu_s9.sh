set -x
mkdir -p gpurun_out/s9
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/s9/pytest.log
timeout 300 python tools/conv_shapes.py --time "32 32 32 192 192 3 0" "32 32 32 192 192 3 1" "32 64 32 192 192 3 0" "32 32 64 192 192 3 1" "32 64 64 96 96 3 0" "32 128 64 96 96 3 0" "32 32 32 192 192 3 2" "32 64 32 192 192 3 2" "32 64 64 96 96 3 2" "32 128 128 48 48 3 2" "32 512 512 12 12 3 2" "32 256 256 24 24 3 2" > gpurun_out/s9/shapes_ksplit1.txt 2>&1
FU_TC_KSPLIT=0 timeout 300 python tools/conv_shapes.py --time "32 32 32 192 192 3 0" "32 32 32 192 192 3 1" "32 64 32 192 192 3 0" "32 32 64 192 192 3 1" "32 64 64 96 96 3 0" "32 128 64 96 96 3 0" > gpurun_out/s9/shapes_ksplit0.txt 2>&1
timeout 300 python tools/step_profile.py > gpurun_out/s9/step_profile.txt 2>&1
(timeout 300 python bench.py 2>&1 | tail -2) > gpurun_out/s9/bench.log
du -sh gpurun_out
