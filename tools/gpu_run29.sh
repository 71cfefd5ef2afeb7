cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# 1. launch list of the bench command (every kernel of 1 warm-up... steps), 2. one full-set capture of one step's kernels
timeout -k 10 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout -k 10 1800 ncu --set full --clock-control none --import-source on -k regex:'scan_mask|scan_bases|scan_expand|aggregate_stream|join_span|join_build|exclusive_scan' -s 24 -c 8 -o gpurun_out/prof_r02_step -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/ncu_full.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out/prof_r02_step.ncu-rep gpurun_out/launches.csv
tail -n 2 gpurun_out/ncu_full.log | cut -c1-300
# 3. the bench lines
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_sf100.json 2> gpurun_out/bench_sf100.err; echo "bench rc=$?"
timeout -k 10 900 python bench.py --sf 10 --steps 10 --warmup 3 > gpurun_out/bench_sf10.json 2> gpurun_out/bench_sf10.err; echo "bench10 rc=$?"
python - <<'PY'
import json
for f in ('bench_sf100','bench_sf10'):
    line=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, {k: line.get(k) for k in ('value','ms_per_step','gpu_launches')}, line['e2e']['value'], line['verify']['ok'])
    print({k:(round(v['kernel_ms'],3), round(v['operator_ms'],3), round(v['frac'],3)) for k,v in line['operators'].items()})
PY
