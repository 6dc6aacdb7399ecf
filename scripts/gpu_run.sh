set -x
PGMP_EDGE_NOGATHER=1 timeout 300 python scripts/quick_profile.py 32 knn tc > gpurun_out/r2_qp_nogather.txt 2>&1; head -4 gpurun_out/r2_qp_nogather.txt
timeout 300 python scripts/quick_profile.py 32 knn tc > gpurun_out/r2_qp_base.txt 2>&1; head -3 gpurun_out/r2_qp_base.txt
