PGMP_LABELS_ST=1 timeout 600 python scripts/time_labels.py 2>&1 | cut -c1-150 | sed -n 2,2p
KMP_BLOCKTIME=0 timeout 600 python scripts/time_labels.py 2>&1 | cut -c1-150 | sed -n 2,2p
OMP_WAIT_POLICY=passive timeout 600 python scripts/time_labels.py 2>&1 | cut -c1-150 | sed -n 2,2p
GOMP_SPINCOUNT=0 timeout 600 python scripts/time_labels.py 2>&1 | cut -c1-150 | sed -n 2,2p
python -c "import torch; print(torch.__config__.parallel_info())" | head -12
