set -x
mkdir -p gpurun_out/s15
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/s15/pytest.log
timeout 300 python tools/step_profile.py > gpurun_out/s15/step_profile.txt 2>&1
(timeout 300 python bench.py 2>gpurun_out/s15/bench.err | tail -2) > gpurun_out/s15/bench.log
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/s15/bench2.err | tail -2) > gpurun_out/s15/bench2.log
du -sh gpurun_out
