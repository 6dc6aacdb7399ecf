set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest21.txt 2>&1; tail -3 gpurun_out/r2_pytest21.txt
python scripts/bench_train.py --steps 5 --warmup 3 > gpurun_out/r2_train4.txt 2>&1; tail -1 gpurun_out/r2_train4.txt | cut -c1-700
