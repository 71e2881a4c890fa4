set -x
mkdir -p gpurun_out/s14
nvidia-smi -L > gpurun_out/s14/gpus.txt
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/s14/bench2.err | tail -2) > gpurun_out/s14/bench2.log
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>gpurun_out/s14/ref2.err | tail -2) > gpurun_out/s14/ref2.log
du -sh gpurun_out
