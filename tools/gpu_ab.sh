# A/B of environment switches: usage: bash tools/gpu_ab.sh <tag> "<ENV=.. ENV=..>" ...   (one bench line + family times per variant)
tag=$1; shift
mkdir -p gpurun_out/$tag
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 200 python bench.py --no-cpu-baseline --no-extras 2>gpurun_out/$tag/bench_$i.err | tail -1 > gpurun_out/$tag/bench_$i.json
  python - <<PY
import json
d=json.load(open("gpurun_out/$tag/bench_$i.json"))
print("[$envs]", round(d["value"],1), "img/s", round(d["ms_per_step"],4), "ms; e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],4))
print("   ", {k:v for k,v in d["engine_ms_by_family"].items() if k in ("act_bwd","bn_fwd","conv3_fwd","conv3_dgrad","conv3_wgrad","_engine_total_ms")})
PY
done
