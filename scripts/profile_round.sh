#!/bin/bash
# Round profile pass (run under gpurun): tests, bench (both arms), ncu launch list, ncu --set full of the two
# kernels the roofline is reported for.  Outputs go to gpurun_out/ and are summarised into profiles/ afterwards.
set -u
R=${1:-r2}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/${R}_pytest_gpu.txt
python bench.py --steps 10 --warmup 3 > $O/${R}_bench.json 2> $O/${R}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_ref.json 2>> $O/${R}_bench.err
python bench.py --steps 2 --warmup 3 > $O/plain_a.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${R}_launches.csv \
      python bench.py --steps 2 --warmup 3 > $O/ncu_a.log 2>&1
python scripts/quick_profile.py 32 knn tc > $O/plain_b.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:edge_step_tc -s 20 -c 1 -o $O/${R}_edge_step_tc \
      python scripts/quick_profile.py 32 knn tc > $O/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nms_candidates -s 2 -c 1 -o $O/${R}_nms \
      python scripts/quick_profile.py 32 knn tc > $O/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:group_kernel -s 1 -c 1 -o $O/${R}_group \
      python scripts/dbg_group.py 32 > $O/ncu_d.log 2>&1
ncu --set full --clock-control none -k regex:gather_features -c 1 -o $O/${R}_gather_features \
      python scripts/quick_profile.py 32 knn tc > $O/ncu_e.log 2>&1
# scoremap assembly fused into the NMS loader: launches 0-12 of nms_candidates are the two-kernel path, 13.. the fused one
python scripts/bench_fused_nms.py > $O/${R}_fused_nms_bench.txt 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:nms_candidates -s 15 -c 1 -o $O/${R}_nms_fused \
      python scripts/bench_fused_nms.py > $O/ncu_f.log 2>&1
cat $O/${R}_pytest_gpu.txt
tail -c 300 $O/${R}_bench.err
