# time the thin 3x3 layer shapes (forward launches with BN statistics) under the environment given on the command line
# usage: bash tools/conv_sweep.sh "<label>" [ENV=VALUE ...]
label=$1; shift
out=$(env "$@" CONV_STATS=1 timeout 100 python tools/conv_shapes.py --time "32 32 32 192 192 3 0" "32 32 32 192 192 3 1" "32 64 64 96 96 3 0" "32 64 64 96 96 3 1" "32 64 32 192 192 3 0" "32 32 64 192 192 3 1" 2>/dev/null | grep " us " | sed 's/.* \([0-9.]*\) us .*/\1/' | tr '\n' ' ')
echo "$label: $out"
