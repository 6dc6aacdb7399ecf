timeout 900 python -m pytest tests -m gpu -x -q -k "mpn or tensor or full or smoke or pipelined" 2>&1 | tail -8
timeout 300 python scripts/quick_profile.py 32 knn tc > gpurun_out/r2_qp_v3b.txt 2>&1; head -8 gpurun_out/r2_qp_v3b.txt
