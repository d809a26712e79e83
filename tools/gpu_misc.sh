#!/bin/bash
# 1-GPU experiments: F1 against box shape (is the config-2 gap per-CTA overhead or segment length?), ingest kernels in HBM
tag=${1:-r2f}
o=gpurun_out/${tag}_roi_shapes.txt; : > $o
for box in 96,65 96,260 96,520 384,65 192,130 288,22; do
  echo "== box $box" >> $o
  timeout 300 python tools/bench_roi.py --frames 4096 --iters 30 --box $box 2>&1 | tail -1 >> $o
done
cat $o
python tools/bench_ingest.py > gpurun_out/${tag}_ingest.txt 2>&1; cat gpurun_out/${tag}_ingest.txt
