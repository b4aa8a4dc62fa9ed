"""Turn a per-launch ncu metrics CSV of tools/one_step.py into profiles/<tag>_kernel_table.md.
    python tools/ncu_table.py gpurun_out/<csv> gpurun_out/plan_names.json profiles/r02_kernel_table.md [event_ms]
The plan names (tools/one_step.py --count) give each launch its op family."""
import collections
import csv
import json
import re
import sys

src, names_path, dst = sys.argv[1:4]
event_ms = sys.argv[4] if len(sys.argv) > 4 else "?"
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
per, order = {}, []
for r in csv.DictReader(lines):
    i = int(r["ID"])
    if i not in per:
        per[i] = {"name": r["Kernel Name"]}
        order.append(i)
    v = r["Metric Value"].replace(",", "")
    try:
        v = float(v)
    except ValueError:
        pass
    per[i][r["Metric Name"]] = v
    per[i]["unit_" + r["Metric Name"]] = r["Metric Unit"]
names = json.load(open(names_path))
assert len(names) == len(order), (len(names), len(order))


def fam(n):
    if re.match(r"b\d+\.", n):
        return "vit." + n.split(".", 1)[1]
    if re.match(r"e\d+\.", n):
        return "extractor." + n.split(".", 1)[1]
    if re.match(r"f\d\.", n):
        return "fapm." + n.split(".", 1)[1]
    if re.match(r"ups\d", n):
        return "ups"
    if n.startswith("tap"):
        return "vit.tap_ln"
    if n.startswith("tail"):
        return "tail"
    if n.startswith("spm.fc"):
        return "spm.fc" + ("1" if n == "spm.fc1" else "2-4")
    return n


T = "gpu__time_duration.sum"
TENS = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
XU = "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"
ISS = "smsp__issue_active.avg.pct_of_peak_sustained_active"
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.OrderedDict()
for i, n in zip(order, names):
    d = per[i]
    a = agg.setdefault(fam(n), dict(n=0, ns=0.0, rd=0.0, wr=0.0, tens=0.0, xu=0.0, issue=0.0, kernel=d["name"],
                                    regs=d.get("launch__registers_per_thread")))
    u = d["unit_" + T]
    t_ns = d[T] * (1e3 if u in ("us", "usecond") else 1e6 if u.startswith("ms") else 1)
    a["n"] += 1
    a["ns"] += t_ns
    a["rd"] += d["dram__bytes_read.sum"] * SCALE[d["unit_dram__bytes_read.sum"]]
    a["wr"] += d["dram__bytes_write.sum"] * SCALE[d["unit_dram__bytes_write.sum"]]
    a["tens"] += d[TENS] * t_ns
    a["xu"] += d[XU] * t_ns
    a["issue"] += d[ISS] * t_ns
tot = sum(a["ns"] for a in agg.values())
peaks = json.load(open("MEASURED_PEAKS.json"))
out = ["# Round 2: every kernel of one dinounet_l forward (B=32, 512x512), ncu per launch\n",
       "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active...,"
       "sm__inst_executed_pipe_xu...,smsp__issue_active... --clock-control none -k regex:<our kernels> -s %d -c %d python "
       "tools/one_step.py dinounet_l 32 512` (%d launches = one eager step after one warm-up step; cold-cache, serialised: "
       "compare shares, not absolutes).  HBM peak = %.0f GB/s (MEASURED_PEAKS.json).  Raw CSV next to this file "
       "(`*_step_metrics.csv`); generator `tools/ncu_table.py`.\n" % (len(order), len(order), len(order), peaks["hbm_gbs"]),
       "| op family (launches) | kernel | regs | total us | share | DRAM rd+wr MB/launch | DRAM GB/s | of HBM peak | tensor pipe % | XU (MUFU) % | issue % |",
       "|---|---|---|---|---|---|---|---|---|---|---|"]
for f, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    us = a["ns"] / 1e3
    mb = (a["rd"] + a["wr"]) / a["n"] / 1e6
    gbs = (a["rd"] + a["wr"]) / a["ns"]
    k = re.sub(r"\(.*", "", a["kernel"]).replace("void ", "")
    out.append(f"| {f} ({a['n']}) | `{k}` | {a['regs']:.0f} | {us:.0f} | {100 * a['ns'] / tot:.1f}% | {mb:.1f} | {gbs:.0f} | "
               f"{gbs / peaks['hbm_gbs']:.2f} | {a['tens'] / a['ns']:.1f} | {a['xu'] / a['ns']:.1f} | {a['issue'] / a['ns']:.1f} |")
out.append(f"\nSum of kernel durations: {tot / 1e6:.2f} ms (CUDA-event step of the same build, graph replay: {event_ms} ms).")
open(dst, "w").write("\n".join(out) + "\n")
# per-launch DRAM traffic of the roofline kernels (bench.py reads this file for its `traffic` fields)
per = {f: (a["rd"] + a["wr"]) / a["n"] for f, a in agg.items()}
vit = [per[k] for k in ("vit.qkv", "vit.proj", "vit.fc1", "vit.fc2") if k in per]
traffic = {"per_launch_dram_bytes": {k: per[k] for k in ("vit.fc1", "vit.fc2", "vit.proj", "vit.qkv", "vit.attn", "d2.conv0", "d2.conv1") if k in per},
           "vit_gemm_family_avg_per_launch": sum(vit) / max(len(vit), 1),
           "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu pass over one eager step with the final binaries (" + src + ")"}
json.dump(traffic, open(dst.replace("kernel_table.md", "traffic.json"), "w"), indent=1)
print("\n".join(out[:16]))
