set -x
mkdir -p gpurun_out/s10
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/s10/pytest.log
timeout 300 python tools/step_profile.py > gpurun_out/s10/step_profile.txt 2>&1
timeout 300 python tools/profile_layers.py 32 192 bf16 > gpurun_out/s10/layers.txt 2>&1
(timeout 300 python bench.py 2>&1 | tail -2) > gpurun_out/s10/bench.log
(timeout 300 python bench.py --batch 8 --size 736 --tile 718 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -2) > gpurun_out/s10/bench_736.log
du -sh gpurun_out
