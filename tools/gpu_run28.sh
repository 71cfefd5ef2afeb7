cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29640 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
python - <<PY
import json
line=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print({k: line.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches')})
print({k:(round(v['kernel_ms'],3), round(v['operator_ms'],3)) for k,v in line['operators'].items()})
print(line['phases_rank0'])
print(line['e2e'])
print(line['verify']['ok'], line['verify'].get('all_ranks_ok'))
PY
tail -n 3 gpurun_out/bench_n$N.err | cut -c1-300
