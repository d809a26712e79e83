#!/bin/bash
# F1 isolated: effect of dirty L2 contents left by the window pipeline, and of the launch size
o=gpurun_out/${1:-r2p}_f1_dirty.txt
: > $o
for d in 0 1 40 80 120; do
  echo "== dirty $d MB, 8192 frames" >> $o
  timeout 300 python tools/bench_roi.py --frames 8192 --iters 40 --dirty $d 2>&1 | tail -1 >> $o
done
echo "== 16384 frames" >> $o
timeout 300 python tools/bench_roi.py --frames 16384 --iters 30 2>&1 | tail -1 >> $o
echo "== 16384 frames dirty 80" >> $o
timeout 300 python tools/bench_roi.py --frames 16384 --iters 30 --dirty 80 2>&1 | tail -1 >> $o
cat $o
