set -x
mkdir -p gpurun_out/s20
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/s20/pytest.log
timeout 200 python tools/step_profile.py > gpurun_out/s20/step_profile.txt 2>&1
FU_TC_FUSE_RES=0 timeout 200 python tools/step_profile.py > gpurun_out/s20/step_profile_nofuse.txt 2>&1
(timeout 300 python bench.py 2>gpurun_out/s20/bench.err | tail -2) > gpurun_out/s20/bench.log
timeout 300 python tools/profile_layers.py 32 192 bf16 > gpurun_out/s20/layers.txt 2>&1
du -sh gpurun_out
