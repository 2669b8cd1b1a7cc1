#!/usr/bin/env python
"""Turn the ncu outputs a gpurun call brought back into the committed summaries under profiles/.

    python tools/ncu_summarize.py <tag> <launches.csv> <full.ncu-rep>

Writes profiles/<tag>_launches.md (per-kernel time shares of one bench.py run),
profiles/<tag>_ncu_full.md (key counters + top stall reasons per kernel) and refreshes
profiles/ncu_traffic.json (dram bytes per launch, read by bench.py for `roofline.traffic`).
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").strip()
    return re.sub(r"^st::", "", name)


# ---- launch list ---------------------------------------------------------------------------------
with open(launches) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
    k = short(row["Kernel Name"])[:90]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
out = [f"# {tag}: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n",
       "`ncu --metrics gpu__time_duration.sum --clock-control none --csv` after the same command exited 0 without ncu.",
       "Times are cold-cache and serialised: compare SHARES with bench.py's live CUDA-event breakdown, not absolutes.\n",
       "| total us | launches | us/launch | share | kernel |", "|---:|---:|---:|---:|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
    out.append(f"| {v[1]:.1f} | {v[0]} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% | `{k}` |")
open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")

# ---- full capture ----------------------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum"]
STALLS = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") or
          (h.startswith("smsp__warp_issue_stalled") and h.endswith("per_warp_active.pct"))]


def to_bytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


traffic = {}
md = [f"# {tag}: `ncu --set full --clock-control none --import-source on` of the hot kernels\n",
      "One launch per kernel from `python bench.py --steps 1 --warmup 3 --no-cpu-baseline` (BASELINE config #3:",
      "2^20 updates + 2^20 samples on 2^27 float32 leaves), captured after the same command exited 0 without ncu.\n"]
seen = set()
for r in rows[2:]:
    name = short(r[idx["Kernel Name"]])
    key = name.split("<")[0]
    if key in seen:
        continue
    seen.add(key)
    md.append(f"\n## `{name}`\n\n| metric | value |\n|---|---|")
    for k in KEEP:
        if k in idx:
            md.append(f"| {k} | {r[idx[k]]} {units[idx[k]]} |")
    st = []
    for h in STALLS:
        try:
            st.append((float(r[idx[h]].replace(",", "")), h))
        except ValueError:
            pass
    st.sort(reverse=True)
    if st:
        md.append("\nTop warp stall reasons (% of active warps per issue slot): " +
                  ", ".join(f"{h.split('stalled_')[1].replace('_per_warp_active.pct', '')} {v:.1f}" for v, h in st[:5]))
    rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
    wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    traffic[key] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                    "source": f"profiles/{tag}_ncu_full.md"}
# the whole update pipeline of one batch (bench.py's roofline.traffic for the update): every update kernel once, the
# radix pass three times (sparse float batches sort 21 key bits), the huge-segment summaries twice (second binade)
UPDATE_KERNELS = {"make_keys_kernel": 1, "cub::DeviceRadixSortOnesweepKernel": 3, "leaf_pass_grouped_kernel": 1, "hy_hist_kernel": 1,
                  "hy_scan_kernel": 1, "hy_scatter_kernel": 1, "hy_segment_kernel": 1, "deep_red_kernel": 1, "retire_kernel": 1,
                  "fold_huge_maps_kernel": 2, "fold_huge_apply_kernel": 1, "fold_scan_kernel": 1}
if all(k in traffic for k in UPDATE_KERNELS):
    tot = sum(traffic[k]["dram_bytes_per_launch"] * c for k, c in UPDATE_KERNELS.items())
    traffic["update_pipeline"] = {"dram_bytes_per_launch": tot, "source": f"profiles/{tag}_ncu_full.md",
                                  "note": "sum over the update kernels of one 2^20-update batch: " +
                                          ", ".join(f"{k} x{c}" for k, c in UPDATE_KERNELS.items())}
    md.append(f"\n## update pipeline, one batch\n\nDRAM traffic summed over its kernels: {tot / 1e6:.1f} MB "
              f"(algorithmic: 236 B x 2^20 = 247.5 MB).")
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full.md"), "w").write("\n".join(md) + "\n")
json.dump({"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full), BASELINE config #3",
           "kernels": traffic}, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print("wrote", tag, sorted(traffic))
