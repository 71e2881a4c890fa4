# Round-2 evidence run on one GPU: whole GPU suite, smoke, the bench line (ours, with parity_tc and other_configs) and the
# reference arm, per-layer profile, and the ncu launch list of one eager step (time, DRAM bytes, tensor pipe, issue slots).
# usage: gpurun --timeout 2400 -- bash tools/gpu_round2.sh ; then tools/ncu_traffic.py on gpurun_out/r02/ncu_launches.csv
set -x
mkdir -p gpurun_out/r02
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r02/pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6) > gpurun_out/r02/smoke.log
(timeout 600 python bench.py 2>gpurun_out/r02/bench.err | tail -1) > gpurun_out/r02/bench.json
(timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1) > gpurun_out/r02/bench_ref.json
timeout 300 python tools/profile_layers.py 32 192 bf16 > gpurun_out/r02/layers.txt 2>&1
FU_STREAMS=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r02/ncu_launches.csv python tools/one_step.py 32 192 2 > gpurun_out/r02/ncu_step.log 2>&1
cat gpurun_out/r02/pytest.log gpurun_out/r02/smoke.log
python -c "
import json; d=json.load(open('gpurun_out/r02/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('parity_tc',{}).get('value')); print([ (o['baseline_config'], round(o['value'],1), o.get('roofline',{}).get('frac')) for o in d.get('other_configs',[])])"
du -sh gpurun_out/r02
