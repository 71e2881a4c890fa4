# Eight-GPU run of the bench (ours): gpurun --gpus 8 -- bash tools/gpu_final_8gpu.sh
mkdir -p gpurun_out/final_8gpu
for n in 8 4; do
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>gpurun_out/final_8gpu/bench$n.err | tail -1) > gpurun_out/final_8gpu/bench$n.json
echo "exit $?" >> gpurun_out/final_8gpu/bench$n.err
done
python - <<'PY'
import json
for f in ("bench8", "bench4"):
    try:
        d = json.load(open(f"gpurun_out/final_8gpu/{f}.json")); print(f, d["n_gpus"], round(d["value"], 1), d["ms_per_step"], round(d["e2e"]["value"], 1))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
tail -2 gpurun_out/final_8gpu/bench8.err
