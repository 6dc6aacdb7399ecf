set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest9.txt 2>&1; tail -3 gpurun_out/r2_pytest9.txt
timeout 300 python scripts/bench_nms.py > gpurun_out/r2_nms_sweep11.txt 2>&1; tail -1 gpurun_out/r2_nms_sweep11.txt | cut -c1-900
