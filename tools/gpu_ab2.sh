#!/bin/bash
# bench.py A/B over discard x batch depth: tools/gpu_ab2.sh <tag>
o=gpurun_out/${1:-r2q}_discard_depth_ab.txt
: > $o
for T in 32 64; do for d in 0 1; do
  echo "== bench.py --frames-per-step $T BPV_DISCARD=$d" >> $o
  BPV_DISCARD=$d timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu --no-other --frames-per-step $T 2>>gpurun_out/ab2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',round(d['value']),'ms/step',round(d['ms_per_step'],4),{k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()},'F1 frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['value']))" >> $o
done; done
tail -5 gpurun_out/ab2.err
cat $o
