#!/bin/bash
# needs a tuning build of the library: BPV_NVCC_EXTRA=-DBPV_ROI_TUNING python -m bpv.build --force  (from bp-from-video_b200/)
# dev aid: time the F1 tuning variants (one process each; BPV_ROI_VARIANT is read once per process)
out=gpurun_out/roi_variants_r1e.txt
: > $out
for v in A10 A11 A12 F12 F14 G08 G12 H12 H24; do
  echo "variant $v" >> $out
  BPV_ROI_VARIANT=$v timeout 300 python tools/bench_roi.py --frames 8192 --iters 30 2>&1 | tail -1 >> $out
done
for box in 32,29 320,160 1920,1080; do
for v in none A12 G12; do
  echo "box $box variant $v" >> $out
  fr=8192; [ $box = 1920,1080 ] && fr=512
  BPV_ROI_VARIANT=$v timeout 300 python tools/bench_roi.py --frames $fr --iters 30 --box $box 2>&1 | tail -1 >> $out
done
done
