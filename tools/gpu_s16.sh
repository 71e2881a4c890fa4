set -x
mkdir -p gpurun_out/s16
(timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/s16/bench2.err | tail -2) > gpurun_out/s16/bench2.log
echo "exit $?" >> gpurun_out/s16/bench2.err
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/s16/pytest.log
timeout 200 python tools/step_profile.py > gpurun_out/s16/step_profile.txt 2>&1
FU_TC_WGRAD_NOEPI=1 timeout 200 python tools/step_profile.py > gpurun_out/s16/step_profile_noepi.txt 2>&1
FU_TC_WGRAD3_WAVES=1 timeout 200 python tools/step_profile.py > gpurun_out/s16/step_profile_w3waves1.txt 2>&1
(timeout 300 python bench.py 2>gpurun_out/s16/bench.err | tail -2) > gpurun_out/s16/bench.log
du -sh gpurun_out
