#!/bin/bash
# 1-GPU experiments: ingest kernels in HBM (and the resize variants)
tag=${1:-r2h}
python -m pytest tests/test_roi_gpu.py -m gpu -q 2>&1 | tail -2
o=gpurun_out/${tag}_ingest.txt; : > $o
python tools/bench_ingest.py >> $o 2>&1
for v in 11 01 00; do BPV_RESIZE_VARIANT=$v python tools/bench_ingest.py 2>&1 | grep resized >> $o; done
cat $o
