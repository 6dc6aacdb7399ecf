set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/bench_train.py --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-560
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_n2b.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','train_step'):
    print(k, json.dumps(d.get(k))[:700])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'grouped', json.dumps(d['grouping_tail'])[:400])
P
