cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -x -q -m gpu > gpurun_out/test_all.log 2>&1; echo "rc=$?" >> gpurun_out/test_all.log
tail -n 4 gpurun_out/test_all.log | cut -c1-400
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
timeout -k 10 300 python tools/variants.py --sf 10 --only aggregate > gpurun_out/variants.txt 2>&1; tail -n 7 gpurun_out/variants.txt
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_sf100.json 2> gpurun_out/bench_sf100.err; echo "bench rc=$?"
timeout -k 10 900 python bench.py --sf 10 --steps 10 --warmup 3 > gpurun_out/bench_sf10.json 2> gpurun_out/bench_sf10.err; echo "bench10 rc=$?"
timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ('bench_sf100','bench_sf10','bench_reference'):
    try:
        line=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, {k: line.get(k) for k in ('value','ms_per_step','gpu_launches')}, line.get('e2e',{}).get('value') if line.get('e2e') else None)
        if 'operators' in line and line['operators']:
            print({k:(round(v['kernel_ms'],3), round(v['operator_ms'],3), round(v['frac'],3)) for k,v in line['operators'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
