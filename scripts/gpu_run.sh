set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "group or pipeline or persons or refine" > gpurun_out/r2_pytest20.txt 2>&1; tail -3 gpurun_out/r2_pytest20.txt
timeout 300 python scripts/quick_profile.py 32 knn tc group > gpurun_out/r2_qp_group2.txt 2>&1; tail -2 gpurun_out/r2_qp_group2.txt
