set -x
mkdir -p gpurun_out/s12
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/s12/pytest.log
(timeout 300 python bench.py 2>gpurun_out/s12/bench.err | tail -2) > gpurun_out/s12/bench.log
(timeout 300 python bench.py --no-graph --no-cpu-baseline 2>&1 | tail -2) > gpurun_out/s12/bench_eager.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/s12/ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-profile > gpurun_out/s12/ncu_bench.log 2>&1
timeout 300 python tools/profile_layers.py 32 192 bf16 > gpurun_out/s12/layers.txt 2>&1
du -sh gpurun_out
