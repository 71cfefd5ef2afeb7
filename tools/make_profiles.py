"""Turn gpurun_out/ scratch captures into the tracked summaries under profiles/.

usage: python tools/make_profiles.py <tag> <step.ncu-rep> <launches.csv> [bench_line.json ...]
Writes profiles/<tag>_kernels.md (per-kernel table from the `ncu --set full` capture), profiles/<tag>_launches.csv (the
`--metrics gpu__time_duration.sum` launch list of `bench.py`), profiles/<tag>_traffic.json (DRAM bytes per launch, read by
bench.py for roofline.traffic) and copies the bench JSON lines.
"""
import csv
import json
import os
import shutil
import subprocess
import sys

tag, report, launches = sys.argv[1], sys.argv[2], sys.argv[3]
extra = sys.argv[4:]
out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
os.makedirs(out_dir, exist_ok=True)

raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
header, units = rows[0], rows[1]


def cell(row, key, scale=1.0):
    if key not in header:
        return None
    text = row[header.index(key)].replace(",", "")
    try:
        value = float(text)
    except ValueError:
        return None
    unit = units[header.index(key)]
    factor = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
              "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}.get(unit, 1.0)
    return value * factor * scale


kernels = {}
for row in rows[2:]:
    name = row[header.index("Kernel Name")].split("(")[0].replace("void ", "").replace("hyb::", "")
    entry = kernels.setdefault(name, [])
    stalls = []
    for i, key in enumerate(header):
        if key.startswith("smsp__average_warps_issue_stalled_") and key.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(row[i]), key[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    entry.append({
        "ms": cell(row, "gpu__time_duration.sum", 1e3),
        "dram_read": cell(row, "dram__bytes_read.sum"),
        "dram_write": cell(row, "dram__bytes_write.sum"),
        "dram_pct": cell(row, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "sm_pct": cell(row, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "occupancy": cell(row, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "regs": cell(row, "launch__registers_per_thread"),
        "grid": cell(row, "launch__grid_size"),
        "block": cell(row, "launch__block_size"),
        "warp_inst": cell(row, "smsp__inst_executed.sum"),
        "stalls": ", ".join(f"{name}={value:.1f}" for value, name in stalls[:4]),
    })

traffic = {}
lines = [f"# {tag}: per-kernel summary of one bench.py step under `ncu --set full --clock-control none`", "",
         "Times under ncu are serialised and cold-cache: use them for shares and for DRAM bytes, not as bench values.", "",
         "| kernel | launches | ms/launch | DRAM read MB | DRAM write MB | DRAM %peak | SM %peak | occupancy % | regs | grid x block | warp inst (M) | top stalls |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for name, entries in kernels.items():
    n = len(entries)
    mean = lambda key: sum(e[key] or 0.0 for e in entries) / n
    traffic[name] = {"dram_bytes_per_launch": mean("dram_read") + mean("dram_write"), "ms_under_ncu": mean("ms")}
    first = entries[0]
    lines.append(f"| {name} | {n} | {mean('ms'):.4f} | {mean('dram_read') / 1e6:.1f} | {mean('dram_write') / 1e6:.1f} | "
                 f"{mean('dram_pct'):.1f} | {mean('sm_pct'):.1f} | {mean('occupancy'):.1f} | {int(first['regs'] or 0)} | "
                 f"{int(first['grid'] or 0)} x {int(first['block'] or 0)} | {mean('warp_inst') / 1e6:.1f} | {first['stalls']} |")
open(os.path.join(out_dir, f"{tag}_kernels.md"), "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(os.path.join(out_dir, f"{tag}_traffic.json"), "w"), indent=1, sort_keys=True)
# bench.py reads the latest capture of its default workload (SF 100, whole job on one GPU) for roofline.traffic
json.dump({"sf": float(os.environ.get("HYB_CAPTURE_SF", "100")), "tag": tag, "kernels": traffic}, open(os.path.join(out_dir, "traffic.json"), "w"), indent=1, sort_keys=True)
shutil.copy(launches, os.path.join(out_dir, f"{tag}_launches.csv"))
for path in extra:
    shutil.copy(path, os.path.join(out_dir, f"{tag}_{os.path.basename(path)}"))
print("\n".join(lines))
