cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 600 python tools/variants.py --sf 10 > gpurun_out/variants.txt 2>&1
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --sf 10 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/ncu_bench.log 2>&1
cat gpurun_out/variants.txt
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches.csv')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
names = rows[hdr]
ki, vi = names.index('Kernel Name'), names.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[hdr + 2:]:
    if len(r) > vi:
        agg[r[ki][:60]].append(float(r[vi].replace(',', '')))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:60s} n={len(v):3d} mean={sum(v)/len(v)/1e3:9.1f} us  last={v[-1]/1e3:9.1f} us")
PY
