# Two-GPU run: the 2-GPU numerical tests, the bench (ours) and a single-GPU bench of the same build for the scaling ratio.
# usage: gpurun --gpus 2 -- bash tools/gpu_final_2gpu.sh
set -x
mkdir -p gpurun_out/final_2gpu
(timeout 600 python -m pytest tests/test_parallel_gpu.py tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -4) > gpurun_out/final_2gpu/pytest.log
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras 2>gpurun_out/final_2gpu/bench2.err | tail -1) > gpurun_out/final_2gpu/bench2.json
echo "exit $?" >> gpurun_out/final_2gpu/bench2.err
(timeout 200 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | tail -1) > gpurun_out/final_2gpu/bench1.json
cat gpurun_out/final_2gpu/pytest.log
python - <<'PY'
import json
for f in ("bench2", "bench1"):
    try:
        d = json.load(open(f"gpurun_out/final_2gpu/{f}.json")); print(f, d["n_gpus"], round(d["value"], 1), d["ms_per_step"], round(d["e2e"]["value"], 1))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
tail -3 gpurun_out/final_2gpu/bench2.err
