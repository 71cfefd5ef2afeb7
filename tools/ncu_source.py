"""Per-source-line instruction / stall-sample shares of one kernel in an .ncu-rep (needs -lineinfo + --import-source)."""
import collections, csv, subprocess, sys
rep, kernel_id = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass,cuda", "--kernel-id", f":::{kernel_id}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
agg = collections.defaultdict(lambda: [0, 0, ""])
cur = None
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    try:
        ln = int(r[hdr.index("Line No")]); n = int(r[hdr.index("Instructions Executed")]); sm = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    key = (cur, ln)
    agg[key][0] += n; agg[key][1] += sm; agg[key][2] = r[1]
tot = sum(v[0] for v in agg.values()) or 1; tots = sum(v[1] for v in agg.values()) or 1
print("total warp instructions", tot, "samples", tots)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(k[0])[:18]:18s}:{k[1]:5d} inst {v[0]/tot*100:5.1f}%  stall {v[1]/tots*100:5.1f}%  {v[2].strip()[:95]}")
