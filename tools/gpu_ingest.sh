#!/bin/bash
tag=${1:-r3a}
python -m pytest tests -m gpu -q -x -k "nv12 or resiz or ingest or roi" 2>&1 | tail -3
python tools/bench_ingest.py 2>&1 | tee gpurun_out/${tag}_ingest.txt
echo "== BPV_NV12_OLD=1 BPV_RESIZE_OLD=1" | tee -a gpurun_out/${tag}_ingest.txt
BPV_NV12_OLD=1 BPV_RESIZE_OLD=1 python tools/bench_ingest.py 2>&1 | grep "NV12\|resized" | tee -a gpurun_out/${tag}_ingest.txt
