SH='32 32 32 192 192 3 0|32 32 32 192 192 3 1|32 64 32 192 192 3 0|32 32 64 192 192 3 1|8 32 32 736 736 3 0|8 32 32 736 736 3 1'
IFS='|' read -ra A <<< "$SH"
for envs in "FU_TC_STACK=0" "FU_TC_STACK=2" "FU_TC_STACK=2 FU_TC_EPI_SETS=1" "FU_TC_STACK=2 FU_TC_STAGING2=0"; do
  echo "== $envs"
  env CONV_STATS=1 FU_TC_VERBOSE=1 $envs timeout 120 python tools/conv_shapes.py --time "${A[@]}" 2>&1 | grep -v "^done\|Warn\|warn" | uniq
done
