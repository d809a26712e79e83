#!/bin/bash
tag=${1:-r3f}
INGEST_FRAMES=4096 timeout 600 ncu --set full --import-source on --clock-control none -k regex:roi_resized_staged --launch-skip 3 --launch-count 1 -f -o gpurun_out/${tag}_resized python tools/bench_ingest.py > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/${tag}_ncu.log
