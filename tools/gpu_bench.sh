# bench.py at SF 100 and SF 10 on one B200 (about 4 GPU-minutes): /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_bench.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_sf100.json 2> gpurun_out/bench_sf100.err; echo "bench rc=$?"
timeout -k 10 900 python bench.py --sf 10 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sf10.json 2> gpurun_out/bench_sf10.err; echo "bench10 rc=$?"
python - <<'PY'
import json
for f in ('bench_sf100','bench_sf10'):
    line=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, {k: line.get(k) for k in ('value','ms_per_step','gpu_launches')}, line['e2e']['value'], line['verify']['ok'])
    print(line['detail']['l2'])
    print({k:(round(v['kernel_ms'],3), round(v['operator_ms'],3), round(v['frac'],3)) for k,v in line['operators'].items()})
PY
