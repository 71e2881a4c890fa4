set -x
mkdir -p gpurun_out/s7
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/s7/pytest.log
timeout 300 python tools/step_profile.py > gpurun_out/s7/step_profile.txt 2>&1
(timeout 300 python bench.py 2>&1 | tail -2) > gpurun_out/s7/bench.log
(timeout 300 python bench.py --torch-loss --no-cpu-baseline 2>&1 | tail -2) > gpurun_out/s7/bench_torchloss.log
du -sh gpurun_out
