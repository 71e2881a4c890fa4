timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for v in 0 1; do echo "== FU_CIN1_MMA=$v"; FU_CIN1_MMA=$v timeout 200 python tools/profile_layers.py 32 192 bf16 2>&1 | grep -i "cin1\|engine kernel"; done
bash tools/gpu_ab.sh cin1 FU_CIN1_MMA=0 FU_CIN1_MMA=1 FU_CIN1_MMA=0 FU_CIN1_MMA=1
