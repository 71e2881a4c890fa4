# A/B of library builds (deepfluorolabeling-ipcai2020_b200/_variants/lib_<name>.so): per-layer profile rows + one bench line each
# usage: bash tools/gpu_libs.sh <tag> <name|default> ...
tag=$1; shift
mkdir -p gpurun_out/$tag
for name in "$@"; do
  if [ "$name" = default ]; then unset FLUORO_UNET_LIB; else export FLUORO_UNET_LIB=$PWD/deepfluorolabeling-ipcai2020_b200/_variants/lib_$name.so; fi
  timeout 200 python tools/profile_layers.py 32 192 bf16 > gpurun_out/$tag/layers_$name.txt 2>&1
  timeout 200 python bench.py --no-cpu-baseline --no-extras 2>gpurun_out/$tag/bench_$name.err | tail -1 > gpurun_out/$tag/bench_$name.json
  python - <<PY
import json
d=json.load(open("gpurun_out/$tag/bench_$name.json"))
print("[$name]", round(d["value"],1), "img/s", round(d["ms_per_step"],4), "ms; e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],4))
PY
done
python - "$tag" "$@" <<'PY'
import sys,re
tag=sys.argv[1]; names=sys.argv[2:]
tabs={}
for n in names:
    t={}
    for line in open(f"gpurun_out/{tag}/layers_{n}.txt"):
        m=re.match(r"(.{34}) (\S+)\s+(\d+)\s+([\d.]+)",line)
        if m: t[(m.group(1).strip(),m.group(2))]=float(m.group(4))
    tabs[n]=t
keys=sorted(tabs[names[0]], key=lambda k:-tabs[names[0]][k])
print("%-34s %-26s"%("tag","kernel")+"".join("%10s"%n for n in names))
for k in keys[:70]:
    print("%-34s %-26s"%k+"".join("%10.4f"%tabs[n].get(k,float('nan')) for n in names))
PY
