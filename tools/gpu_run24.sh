cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -x -q -m gpu > gpurun_out/test_all.log 2>&1; echo "rc=$?" >> gpurun_out/test_all.log
tail -n 8 gpurun_out/test_all.log | cut -c1-600
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_sf100.json 2> gpurun_out/bench_sf100.err; echo "rc=$?"
python - <<'PY'
import json
line=json.loads(open('gpurun_out/bench_sf100.json').read().strip().splitlines()[-1])
print({k: line[k] for k in ('value','ms_per_step','gpu_launches')})
print(line['operators'])
print(line['roofline'])
print(line['e2e'])
print(line['cpu_baseline'])
print(line['verify']['ok'])
PY
tail -n 3 gpurun_out/bench_sf100.err | cut -c1-300
