"""Per-source-line stall samples of one kernel in an .ncu-rep, sorted by samples, with the dominant stall reasons."""
import collections, csv, subprocess, sys
rep, kernel_id = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass,cuda", "--kernel-id", f":::{kernel_id}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; cur = None
agg = collections.defaultdict(lambda: [0, collections.Counter(), ""])
total_reasons = collections.Counter()
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    try:
        ln = int(r[hdr.index("Line No")]); sm = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    key = (cur, ln); agg[key][0] += sm; agg[key][2] = r[1]
    for i, name in enumerate(hdr):
        if name.startswith("stall_") and "Not Issued" not in name:
            try: v = int(r[i])
            except ValueError: continue
            agg[key][1][name[6:]] += v; total_reasons[name[6:]] += v
tot = sum(v[0] for v in agg.values()) or 1
print("samples", tot, "reasons:", ", ".join(f"{k}={v/tot*100:.1f}%" for k, v in total_reasons.most_common(8)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    reasons = ", ".join(f"{n}={c}" for n, c in v[1].most_common(3))
    print(f"{str(k[0])[:16]:16s}:{k[1]:5d} {v[0]/tot*100:5.1f}%  [{reasons}]  {v[2].strip()[:80]}")
