#!/bin/bash
# F1 A/B on the GPU box: one-ROI-per-CTA staged kernel (0) against the persistent double-buffered kernel (1)
o=gpurun_out/${1:-r2d}_roi_ab.txt
: > $o
for v in 0 1; do
  echo "== BPV_ROI_PIPELINED=$v, config-2 boxes, 8192 frames" >> $o
  BPV_ROI_PIPELINED=$v timeout 300 python tools/bench_roi.py --frames 8192 --iters 40 2>&1 | tail -2 >> $o
done
for box in 32,29 320,160 640,360; do
for v in 0 1; do
  echo "== BPV_ROI_PIPELINED=$v box $box" >> $o
  BPV_ROI_PIPELINED=$v timeout 300 python tools/bench_roi.py --frames 4096 --iters 30 --box $box 2>&1 | tail -1 >> $o
done
done
for v in 0 1; do
  echo "== bench.py BPV_ROI_PIPELINED=$v" >> $o
  BPV_ROI_PIPELINED=$v timeout 600 python bench.py --steps 60 --warmup 5 --no-cpu --no-other 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step',d['ms_per_step'],'roi us',d['kernels']['roi']['ms']*1e3,'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])" >> $o
done
cat $o
