# Two-GPU run of the bench (ours with and without graph replay, reference arm): gpurun --gpus 2 -- bash tools/gpu_final_2gpu.sh
set -x
mkdir -p gpurun_out/final_2gpu
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/final_2gpu/bench2.err | tail -2) > gpurun_out/final_2gpu/bench2.log
echo "exit $?" >> gpurun_out/final_2gpu/bench2.err
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-graph 2>/dev/null | tail -1) > gpurun_out/final_2gpu/bench2_eager.log
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1) > gpurun_out/final_2gpu/ref2.log
