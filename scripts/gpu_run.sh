set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest17.txt 2>&1; tail -3 gpurun_out/r2_pytest17.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; tail -c 400 gpurun_out/r2_bench3.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench3.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','serial_calls','e2e','roofline','roofline_nms','kernels','gpu_launches'):
    print(k, json.dumps(d.get(k))[:900])
P
