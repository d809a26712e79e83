#!/bin/bash
# bench.py A/B over an environment switch: tools/gpu_ab.sh VAR v1 v2 ...
var=$1; shift
for v in "$@"; do
  echo "== bench.py $var=$v"
  env $var=$v timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu --no-other 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step',round(d['ms_per_step'],4),{k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()},'F1 frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['value']))"
done
