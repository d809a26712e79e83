#!/bin/bash
# needs a tuning build of the library: BPV_NVCC_EXTRA=-DBPV_ROI_TUNING python -m bpv.build --force  (from bp-from-video_b200/)
# dev aid: F1 reading ROI rows zero-copy from pinned host memory over PCIe, per tuning variant
out=gpurun_out/roi_host_variants_r1e.txt
: > $out
for v in none 544 044 244 G12 B24 D24 E20; do
  echo "host variant $v" >> $out
  BPV_ROI_VARIANT=$v timeout 300 python tools/bench_roi.py --frames 1024 --iters 10 --host 2>&1 | tail -1 >> $out
done
