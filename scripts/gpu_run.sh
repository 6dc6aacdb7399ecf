export PGMP_NVCC_EXTRA=-DPGMP_TIMELINE
timeout 900 python -m pytest tests -m gpu -x -q -k "mpn or tensor or full or smoke or pipelined" 2>&1 | tail -3
timeout 300 python scripts/quick_profile.py 32 knn tc 2>&1 | sed -n 1,4p
timeout 300 python scripts/step_timeline.py 2>&1 | grep -A8 "CTA 3 tile group 0"
PGMP_STEP_ONE_GROUP=1 timeout 300 python scripts/quick_profile.py 32 knn tc 2>&1 | sed -n 2,2p
PGMP_STEP_ONE_GROUP=1 timeout 300 python scripts/step_timeline.py 2>&1 | grep -A8 "CTA 3 tile group 0"
