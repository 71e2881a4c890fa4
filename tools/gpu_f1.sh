mkdir -p gpurun_out/f1
timeout 900 python -m pytest tests/test_loss_gpu.py -x -q -m gpu 2>&1 | tail -3
STEP_IN_HEADS=0 timeout 200 python tools/step_profile.py 2>&1 | grep -i "loss\|# "
for fl in "--separate-loss" "--separate-loss"; do
  timeout 200 python bench.py --no-cpu-baseline --no-extras $fl 2>gpurun_out/f1/bench.err | tail -1 > gpurun_out/f1/bench.json
  python - <<PY
import json
d=json.load(open("gpurun_out/f1/bench.json"))
print("[$fl]", round(d["value"],1), "img/s", round(d["ms_per_step"],4), "ms; e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],4), "launches", d["gpu_launches"])
PY
done
