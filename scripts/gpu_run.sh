timeout 900 python -m pytest tests/test_gpu_train.py -q -k "fixture or full_size" 2>&1 | grep -E "^E  |Error|assert|passed|failed" | head -40
