set -x
for i in 1 2 3; do timeout 900 python -m pytest tests -m gpu -x -q -k "group or pipeline or persons or refine" 2>&1 | tail -1; done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest23.txt 2>&1; tail -2 gpurun_out/r2_pytest23.txt
