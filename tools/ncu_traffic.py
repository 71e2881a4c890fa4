"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
one line per kernel (launches, time, DRAM bytes per training step) + a compact per-launch CSV.
usage: ncu_traffic.py <ncu.csv> <out_launches.csv> <out_summary.json>"""
import csv, json, re, sys, collections
src, out_csv, out_json = sys.argv[1:4]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
launch = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) != len(hdr):
        continue
    L = launch.setdefault(int(r[col["ID"]]), {"kernel": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
    v = float(r[col["Metric Value"]].replace(",", ""))
    unit = r[col["Metric Unit"]]
    name = r[col["Metric Name"]]
    if name.startswith("gpu__time_duration"):
        L["us"] = v / 1e3 if unit.startswith("ns") or unit == "nsecond" else (v if unit.startswith("us") else v * 1e3)
    elif name.startswith("dram__bytes_read"):
        L["rd"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    elif name.startswith("dram__bytes_write"):
        L["wr"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    elif name.startswith("sm__pipe_tensor_cycles_active"):
        L["tensor"] = v
    elif name.startswith("smsp__issue_active"):
        L["issue"] = v
def short(k):
    k = re.sub(r"\(.*", "", k).replace("void ", "").replace("fu::", "")
    return re.sub(r"<__nv_bfloat16(, )?", "<", k)
with open(out_csv, "w") as f:
    f.write("id,kernel,grid,block,time_us,dram_read_bytes,dram_write_bytes,tensor_pipe_pct,issue_active_pct\n")
    for i, L in launch.items():
        f.write(f"{i},\"{short(L['kernel'])}\",\"{L['grid']}\",\"{L['block']}\",{L.get('us', 0):.2f},{int(L.get('rd', 0))},{int(L.get('wr', 0))},"
                f"{L.get('tensor', 0):.1f},{L.get('issue', 0):.1f}\n")
# training steps in the capture = launches of a kernel that runs exactly once per step
steps = (sum(1 for L in launch.values() if "nchw_to_nhwc_kernel" in L["kernel"]) or
         sum(1 for L in launch.values() if "heads_bwd_fused_kernel" in L["kernel"]) or 1)
agg = collections.OrderedDict()
for L in launch.values():
    a = agg.setdefault(short(L["kernel"]), {"launches": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
    a["launches"] += 1; a["us"] += L.get("us", 0); a["rd"] += L.get("rd", 0); a["wr"] += L.get("wr", 0)
summ = {"steps_captured": steps, "per_step": {}}
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    summ["per_step"][k] = {"launches": round(a["launches"] / steps, 2), "ms": round(a["us"] / steps / 1e3, 4),
                           "dram_read_MB": round(a["rd"] / steps / 1e6, 2), "dram_write_MB": round(a["wr"] / steps / 1e6, 2)}
tc = [v for k, v in summ["per_step"].items() if k.startswith("tc_conv") or k.startswith("tc_wgrad")]
summ["tensor_core_conv_kernels_per_step"] = {"ms": round(sum(v["ms"] for v in tc), 4),
                                             "dram_bytes": int(sum(v["dram_read_MB"] + v["dram_write_MB"] for v in tc) * 1e6)}
summ["all_kernels_per_step"] = {"ms": round(sum(v["ms"] for v in summ["per_step"].values()), 4),
                                "dram_bytes": int(sum(v["dram_read_MB"] + v["dram_write_MB"] for v in summ["per_step"].values()) * 1e6)}
json.dump(summ, open(out_json, "w"), indent=1)
print(json.dumps(summ["tensor_core_conv_kernels_per_step"]), json.dumps(summ["all_kernels_per_step"]), "steps", steps)
for k, v in list(summ["per_step"].items())[:16]:
    print(f"{k:44s} {v}")
