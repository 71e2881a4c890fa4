"""Text summary of one `ncu --set full --import-source on` capture (run here, no GPU needed):
key metrics of the launch, where the executed instructions and the stall samples go by code region, and the most
sampled SASS instructions.  usage: ncu_summary.py <file.ncu-rep> [> profiles/xxx.txt]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]


def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr, vals = raw[0], raw[-1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
print(f"# {rep}")
for h, v in zip(hdr, vals):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
        print(f"{h:86s} {v}")
src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--print-source", "sass"]))))
h2 = src[1]
data = src[2:]
ia, isamp, isrc = h2.index("Instructions Executed"), h2.index("# Samples"), h2.index("Source")
reasons = [c for c in h2 if c.startswith("stall_") and "Not Issued" not in c]
ridx = {r: h2.index(r) for r in reasons}
tot_i = sum(int(r[ia]) for r in data)
tot_s = sum(int(r[isamp]) for r in data)
print(f"\n# source page: {len(data)} SASS instructions, {tot_i} warp instructions executed, {tot_s} stall samples")
print("# by code region (100 SASS instructions each; regions with >= 1 % of the samples)")
marks = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "SYNCS", "BAR.SYNC", "ATOM", "RED", "LDS", "STS", "LDG", "STG", "UTCBAR", "UBLKRED")
for s in range(0, len(data), 100):
    seg = data[s:s + 100]
    n = sum(int(r[ia]) for r in seg)
    sm = sum(int(r[isamp]) for r in seg)
    if sm < 0.01 * tot_s:
        continue
    kinds = sorted({k for r in seg for k in marks if k in r[isrc]})
    d = {k: sum(int(r[i] or 0) for r in seg) for k, i in ridx.items()}
    top = [(k[6:], v) for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:4] if v]
    print(f"  sass[{s:5d}:{s + 100:5d}] instr {100 * n / tot_i:5.1f}%  samples {100 * sm / tot_s:5.1f}%  {','.join(kinds):40s} {top}")
mm = [i for i, r in enumerate(data) if "UTCHMMA" in r[isrc]]
if mm:
    a, b = max(0, mm[0] - 450), min(len(data), mm[-1] + 60)
    d = {k: sum(int(r[i] or 0) for r in data[a:b]) for k, i in ridx.items()}
    t = sum(d.values())
    print(f"\n# MMA-issuing warp (sass[{a}:{b}]): {t} samples: " + ", ".join(f"{k[6:]} {v}" for k, v in sorted(d.items(), key=lambda kv: -kv[1]) if v))
print("\n# most sampled instructions")
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:16]):
    print(f"  {i:5d} {data[i][isrc].strip()[:72]:72s} executed {data[i][ia]:>9s} samples {data[i][isamp]:>5s}")
