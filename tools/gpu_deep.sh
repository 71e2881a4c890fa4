SH='32 512 1024 6 6 3 1|32 1024 1024 6 6 3 1|32 512 1024 6 6 3 0|32 1024 1024 6 6 3 0|32 1024 512 12 12 3 0|32 512 1024 12 12 3 1'
IFS='|' read -ra A <<< "$SH"
for envs in "FU_X=1" "FU_TC_MC=1" "FU_TC_BN_MAX=64" "FU_TC_BN_MAX=128"; do
  echo "== $envs"
  env CONV_STATS=1 $envs timeout 120 python tools/conv_shapes.py --time "${A[@]}" 2>&1 | grep -v "^done\|Warn\|warn"
done
