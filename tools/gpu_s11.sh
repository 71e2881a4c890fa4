set -x
mkdir -p gpurun_out/s11
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/s11/pytest.log
(timeout 300 python bench.py 2>gpurun_out/s11/bench.err | tail -2) > gpurun_out/s11/bench.log
(timeout 300 python bench.py --no-graph --no-cpu-baseline 2>&1 | tail -2) > gpurun_out/s11/bench_eager.log
du -sh gpurun_out
