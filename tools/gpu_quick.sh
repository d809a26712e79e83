#!/bin/bash
# quick GPU check: selected tests + bench family times (serial F3/F4 and default schedule)
tag=${1:-r2r}; sel=${2:-xcorr}
python -m pytest tests -m gpu -q -x -k "$sel" 2>&1 | tail -5
for ov in 1 3; do
  echo "== BPV_OVERLAP=$ov"
  BPV_OVERLAP=$ov timeout 900 python bench.py --steps 60 --warmup 5 --no-cpu --no-other 2>>gpurun_out/${tag}_quick.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',round(d['value']),'ms/step',round(d['ms_per_step'],4),{k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()},'F1 frac',round(d['roofline']['frac'],4),'F2 frac',round(d['roofline_by_time']['frac'],3),'e2e',round(d['e2e']['value']))"
done
tail -3 gpurun_out/${tag}_quick.err
