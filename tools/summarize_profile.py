"""Turn gpurun_out/*.ncu-rep + launches csv into the small, tracked summaries under profiles/.
usage: python tools/summarize_profile.py TAG   (reads gpurun_out/{prof_corr,prof_spec}_TAG.ncu-rep, launches_TAG.csv, bench_TAG.json)"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    return [{"name": r[h.index("Kernel Name")], **{k: (r[h.index(k)], units[h.index(k)]) for k in KEYS if k in h},
             **{n: (r[i], units[i]) for i, n in enumerate(h) if "issue_stalled" in n and n.endswith("per_issue_active.ratio")}}
            for r in rows[2:]]


summary = {"tag": tag, "kernels": {}}
lines = [f"# ncu summary {tag}", ""]
for short in ("corr", "spec"):
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{short}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    recs = raw(rep)
    r = recs[-1]
    lines += [f"## {r['name']}  (ncu --set full --clock-control none; last of {len(recs)} captured launches)", "", "| metric | value | unit |", "|---|---|---|"]
    for k, v in r.items():
        if k != "name":
            lines.append(f"| {k.replace('smsp__average_warps_issue_stalled_', 'stall: ')} | {v[0]} | {v[1]} |")
    lines.append("")
    def num(key):
        v, u = r.get(key, ("0", ""))
        x = float(v.replace(",", ""))
        return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
    summary["kernels"][r["name"].split("(")[0]] = {"dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                                                   "duration_ns_under_ncu": r.get("gpu__time_duration.sum", ("", ""))[0]}
lp = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    d = collections.defaultdict(list)
    for r in rows[hdr + 1:]:
        if len(r) == len(h):
            rec = dict(zip(h, r))
            d[rec["Kernel Name"].split("(")[0]].append(float(rec["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    lines += ["## launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare shares)", "",
              "| kernel | launches | mean ns | share of GPU time |", "|---|---|---|---|"]
    for k, v in d.items():
        lines.append(f"| {k} | {len(v)} | {sum(v) / len(v):.0f} | {100 * sum(v) / tot:.1f}% |")
        summary.setdefault("launch_list", {})[k] = {"launches": len(v), "mean_ns": sum(v) / len(v), "share": sum(v) / tot}
    with open(os.path.join(out_dir, f"launches_{tag}.csv"), "w") as f:
        f.write(open(lp).read())
bp = os.path.join(ROOT, "gpurun_out", f"bench_{tag}.json")
if os.path.exists(bp) and os.path.getsize(bp):
    summary["bench_line"] = json.loads(open(bp).read().strip().splitlines()[-1])
    lines += ["", "## bench line of the same build (not under a profiler)", "", "```", json.dumps(summary["bench_line"]), "```"]
open(os.path.join(out_dir, f"ncu_{tag}.md"), "w").write("\n".join(lines) + "\n")
json.dump(summary, open(os.path.join(out_dir, f"summary_{tag}.json"), "w"), indent=1)
k = summary["kernels"].get("void gb::k_correlate_cells<8, 2, 0>") or next((v for n, v in summary["kernels"].items() if "correlate" in n), None)
if k:  # printed for the hand-maintained profiles/traffic.json (which also keeps the other launch sizes and says how it was taken)
    print("correlate dram bytes per launch:", k["dram_bytes_per_launch"])
print("\n".join(lines[:60]))
