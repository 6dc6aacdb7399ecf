set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "nms or graph_constructor or no_threshold or capacities or full_size or pipeline" > gpurun_out/r2_pytest11.txt 2>&1; tail -3 gpurun_out/r2_pytest11.txt
timeout 300 python scripts/bench_nms.py > gpurun_out/r2_nms_sweep13.txt 2>&1; tail -1 gpurun_out/r2_nms_sweep13.txt | cut -c1-900
