# One gpurun call that re-validates the round on a single B200 (about 3 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/gpu_validate.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -x -q -m gpu > gpurun_out/test_all.log 2>&1; echo "rc=$?" >> gpurun_out/test_all.log
tail -n 4 gpurun_out/test_all.log | cut -c1-400
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/smoke.log 2>&1; tail -n 3 gpurun_out/smoke.log | cut -c1-300
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_sf100.json 2> gpurun_out/bench_sf100.err; echo "bench rc=$?"
python - <<'PY'
import json
line = json.loads(open('gpurun_out/bench_sf100.json').read().strip().splitlines()[-1])
print({k: line.get(k) for k in ('value', 'ms_per_step', 'gpu_launches')}, line['e2e']['value'], line['verify']['ok'], line['roofline']['traffic'])
print({k: (round(v['kernel_ms'], 3), round(v['operator_ms'], 3), round(v['frac'], 3)) for k, v in line['operators'].items()})
PY
