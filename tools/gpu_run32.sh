cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo skip
timeout -k 10 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sf100.json 2> gpurun_out/bench_sf100.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('bench_sf10','bench_sf100'):
    try:
        line=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, {k: line.get(k) for k in ('value','ms_per_step')}, line['e2e'], line['verify']['ok'])
    except Exception as e:
        print(f, 'ERR', e); print(open(f'gpurun_out/{f}.err').read()[-1500:])
PY
